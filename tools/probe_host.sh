#!/bin/bash
# what the host side of a GPU box looks like: cores, NUMA, GPU affinity, memory and PCIe bandwidth
cd /root/repo
{
echo "== nproc: $(nproc)   cpuset: $(cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null)"
lscpu | egrep -i 'model name|socket|numa|^cpu\(s\)|thread|core|L3|L2'
echo "== nodes"; ls /sys/devices/system/node/ | grep node; for n in /sys/devices/system/node/node*; do echo "$n: cpus $(cat $n/cpulist) mem $(grep MemTotal $n/meminfo | awk '{print $4}') kB"; done
echo "== taskset: $(taskset -p $$)"
echo "== mems allowed: $(grep -i mems_allowed_list /proc/self/status)"
free -g | head -2
nvidia-smi topo -m
nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist)"; fi; done
which numactl; ulimit -l
echo "== hostbw"; build/hostbw 256
echo "== pcie"; python tools/pcie_bw.py
} > gpurun_out/probe_host.txt 2>&1
