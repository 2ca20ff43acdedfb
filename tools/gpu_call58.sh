#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -k regex:"stem_im2col|conv_tc4|conv_wgrad2|pad_rows|tf32_split" -c 12 --csv --log-file gpurun_out/c58_stem.csv python tools/prof_stem.py > gpurun_out/c58.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/c58_stem.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size'); ii=hdr.index('ID')
cur={}
for r in rows[start+1:]:
    if len(r)<=vi: continue
    cur.setdefault((r[ii], r[ki][:40], r[gi]), {})[r[mi]]=r[vi]
for k,v in cur.items(): print(k, v)
PY
