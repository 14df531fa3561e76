python -m pytest tests -m gpu -x -q -k "rollout or cfg4" 2>&1 | tail -2
for B in 28416 37888 40000 65536 131072; do for K in 1 3; do echo "knob $K B $B: $(MPK_FD_SPLIT=$K python scripts/fd_probe.py $B 1000 3 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_min"])')"; done; done
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/fd_probe.py 40000 6 1 2>&1 | tail -1
