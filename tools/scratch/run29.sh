timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_msm_merge|k_msm_accumulate|k_msm_item_counts" -c 400 --csv --log-file gpurun_out/merge_list.csv python tools/msm_bench.py 20 22 > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/merge_list.csv') if not l.startswith('=='))]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
seq=[(r[ki].split('(')[0], r[gi], float(r[vi])/1e3) for r in rows[1:]]
# print every 7th repetition pattern: show the last rep of each case
for i,(k,g,v) in enumerate(seq):
    print(i, k, g, round(v,1))
P
