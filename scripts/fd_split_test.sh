python -m pytest tests -m gpu -x -q -k "rollout" 2>&1 | tail -2
for K in 3; do for B in 2048 4736 8192 9472; do echo "knob $K B $B: $(MPK_FD_SPLIT=$K python scripts/fd_probe.py $B 1000 3 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_min"], d["checksum"])')"; done; done
