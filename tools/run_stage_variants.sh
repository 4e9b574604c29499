step() { python bench.py --steps 100 --no-cpu --no-e2e --no-multi --no-hard "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step']*1e3/12,2),'us/scan  decode',round(d['roofline']['us_per_launch'],2),'iso',round(d['roofline']['isolated']['us_per_launch'],2))"; }
echo "default  : $(step)"
for f in 3dscanner-graycode_b200/lib/variants/*STAGES*.so; do
  echo "$(basename $f): $(SLB200_LIB=$PWD/$f step)  reserve 0: $(SLB200_LIB=$PWD/$f step --sm-reserve 0)"
done
