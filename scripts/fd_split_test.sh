for L in manipulapy_b200/_lib manipulapy_b200/_lib_nofmad; do
echo "== $L"
MPK_LIB_DIR=$L python scripts/fd_bits2.py 1000 | python -c "
import json,sys; d=json.load(sys.stdin)
for k,v in d.items(): print(k, {a:(b['differing'], b.get('max_ulps')) for a,b in v.items()})"
for B in 8192 65536; do echo "B $B: $(MPK_LIB_DIR=$L python scripts/fd_probe.py $B 1000 3 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_min"])')"; done
done
