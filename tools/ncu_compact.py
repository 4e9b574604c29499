"""ncu --csv launch list (one row per launch and metric) -> the compact form kept under profiles/:
   id,kernel,grid,block,metric,value,unit      usage: python tools/ncu_compact.py raw.csv "comment" > profiles/x.csv"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
col = {n: h.index(n) for n in ("ID", "Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Value", "Metric Unit")}
for line in sys.argv[2:]:
    print("# " + line)
print("id,kernel,grid,block,metric,value,unit")
w = csv.writer(sys.stdout, lineterminator="\n")
for r in rows[1:]:
    name = r[col["Kernel Name"]].split("(")[0]
    w.writerow([r[col["ID"]], name, r[col["Grid Size"]], r[col["Block Size"]], r[col["Metric Name"]],
                r[col["Metric Value"]].replace(",", ""), r[col["Metric Unit"]]])
