# A/B of library build variants: parity tests + the bench step + batch; SLB200_LIB selects the build
V=3dscanner-graycode_b200/lib/variants
run() { python bench.py --steps 100 --no-cpu --no-e2e --no-hard "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['us_per_scan'],2),'us/scan  decode',round(d['roofline']['us_per_launch'],2),'off',round(d['stage2']['offsets_us'],2),'emit',round(d['stage2']['emit_tables_us'],2), 'batch64 ms', round(d['multi']['config3_64_views_by_scan']['compute_only']['ms_per_step'],3), '4K us', round(d['multi']['config2_4k_row_bands']['compute_only']['ms_per_step']*1e3,1), 'parity', d.get('parity_checked'))"; }
echo "default: $(run)"
for f in $V/*.so; do
  echo "== $(basename $f)"
  SLB200_LIB=$PWD/$f python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary_r2.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -1
  echo "   $(SLB200_LIB=$PWD/$f run)"; echo "   $(SLB200_LIB=$PWD/$f run)"
done
