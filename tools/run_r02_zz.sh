#!/bin/bash
cd /root/repo
python -m pytest tests -q -m gpu > gpurun_out/r02zz_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02zz_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
