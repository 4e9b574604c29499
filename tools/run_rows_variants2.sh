# pipelined bench step: stream / graph shapes and emit_rows CTA shapes
step() { python bench.py --steps 100 --no-cpu --no-e2e --no-multi --no-hard "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step']*1e3/12,2),'us/scan  emit_rows',round(d['stage2']['emit_rows_us'],2))"; }
echo "default            : $(step)"
echo "streams 3          : $(step --streams 3)"
echo "streams 6          : $(step --streams 6)"
echo "scans-per-graph 2  : $(step --scans-per-graph 2)"
echo "scans-per-graph 3  : $(step --scans-per-graph 3)"
for f in 3dscanner-graycode_b200/lib/variants/*ROWS_WARPS*.so; do
  echo "$(basename $f): $(SLB200_LIB=$PWD/$f step)  reserve 4: $(SLB200_LIB=$PWD/$f step --sm-reserve 4)"
done
