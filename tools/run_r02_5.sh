#!/bin/bash
cd /root/repo
python -m pytest tests/test_gpu_single.py -x -q -m gpu > gpurun_out/r02_5_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_5_tests.log
python tools/latency_probe.py > gpurun_out/r02_latency_probe.txt 2>&1; cat gpurun_out/r02_latency_probe.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_verify_small -c 20 --csv --log-file gpurun_out/r02_small_launches.csv python tools/latency_probe.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_small_launches.csv', errors='replace')) if len(r)>10]
ix={n:i for i,n in enumerate(rows[0])}
v=[float(r[ix['Metric Value']].replace(',','')) for r in rows[1:] if r[ix['Metric Name']]=='gpu__time_duration.sum']
print('k_verify_small device time under ncu (ns):', v[:20], rows[1][ix['Metric Unit']] if len(rows)>1 else '')
PY
