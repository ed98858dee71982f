#!/bin/bash
cd /root/repo
python tools/hybrid_probe.py 1000000 > gpurun_out/r02_hybrid_probe.txt 2> gpurun_out/r02_hybrid_probe.err; echo "rc=$?"; cat gpurun_out/r02_hybrid_probe.txt; tail -3 gpurun_out/r02_hybrid_probe.err
