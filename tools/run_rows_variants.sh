# A/B of emit_rows_kernel build variants (tools/build_variants.sh "SL_ROWS_..."): isolated and batch emission, then the
# pipelined bench step with the row-mapped emission
V=3dscanner-graycode_b200/lib/variants
echo "default: $(python tools/bench_emit_rows.py 2>&1 | tail -2 | tr '\n' ' ')"
for f in $V/*ROWS*.so; do
  echo "$(basename $f): $(SLB200_LIB=$PWD/$f python tools/bench_emit_rows.py 2>&1 | tail -2 | tr '\n' ' ')"
done
step() { python bench.py --steps 100 --no-cpu --no-e2e --no-multi --no-hard 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step']*1e3/12,2),'us/scan  decode',round(d['roofline']['us_per_launch'],2),'off',round(d['stage2']['offsets_us'],2),'emit',round(d['stage2']['emit_tables_us'],2),'parity',d.get('parity_checked'))"; }
echo "step tiles: $(step)"
echo "step rows : $(SLB200_EMIT=rows step)"
echo "step rows : $(SLB200_EMIT=rows step)"
