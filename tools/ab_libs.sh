# Same-box A/B of library builds: bash tools/ab_libs.sh O A B   (build/lib_<V>/libinfur_b200.so; B = the tree's own build)
set -u
cp infur_b200/lib/libinfur_b200.so build/lib_cur.so
mkdir -p build/lib_B && cp build/lib_cur.so build/lib_B/libinfur_b200.so
for round in 1 2; do
  for v in "$@"; do
    cp build/lib_$v/libinfur_b200.so infur_b200/lib/libinfur_b200.so
    python bench.py --model f16 --no-cpu-baseline --min-seconds 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v f16  round $round', round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))"
    python bench.py --model int8 --no-cpu-baseline --min-seconds 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); i=d.get('int8',d); print('$v int8 round $round', round(i['value'],1), round(i['e2e']['value'],1))"
  done
done
cp build/lib_cur.so infur_b200/lib/libinfur_b200.so
