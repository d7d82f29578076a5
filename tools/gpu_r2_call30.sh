#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_convolution.py -m gpu -q -x 2>&1 | tail -2
python tools/c4_conv_time.py
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:convolve -c 2000 --csv --log-file gpurun_out/c30_conv.csv python tools/c4_conv_time.py > gpurun_out/c30_ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/c30_conv.csv')) if len(r)>5]
h=rows[0]; i=h.index('Metric Value'); u=h.index('Metric Unit')
vals=[float(r[i].replace(',',''))*({'nsecond':1e-6,'ns':1e-6,'usecond':1e-3,'us':1e-3,'msecond':1,'ms':1}.get(r[u],1e-6)) for r in rows[1:]]
print(len(vals),'convolve launches, total ms', sum(vals), 'second half', sum(vals[len(vals)//2:]), 'max', max(vals))
PY
