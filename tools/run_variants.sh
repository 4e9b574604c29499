# A/B of library build variants: parity tests, the bench step, emit beside a decode chain; SLB200_LIB selects the build
V=3dscanner-graycode_b200/lib/variants
run() { python bench.py --steps 100 --no-multi --no-cpu --no-e2e --no-hard "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['us_per_scan'],2),'us/scan  decode',round(d['roofline']['us_per_launch'],2),'off',round(d['stage2']['offsets_us'],2),'emit',round(d['stage2']['emit_tables_us'],2))"; }
echo "default: $(run)"; python tools/bench_overlap2.py | tail -1
for f in $V/*.so; do
  echo "== $(basename $f)"
  SLB200_LIB=$PWD/$f python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary_r2.py -m gpu -x -q 2>&1 | tail -1
  for r in 8 0; do echo "   reserve $r: $(SLB200_LIB=$PWD/$f run --sm-reserve $r)"; done
  SLB200_LIB=$PWD/$f python tools/bench_overlap2.py
done
