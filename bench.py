#!/usr/bin/env python
"""bench.py -- MPT proofs verified / s (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path on the host cores

Workload at every N (weak scaling, one process per GPU, no data-path collective): BASELINE.json
configs[1] -- a batch of 1 M account proofs against a synthetic 10 M-account state trie (SURVEY.md
section 8d config 2), built by workload/ (no RPC, no network).  A "step" is one pass of the whole
verification hot path (K0 binning, K1 Keccak-256 of every node, K2a decode, K2b walk) over that
batch.  The 3 GB node arena is far larger than L2 (126 MB), so every step streams from HBM.

`value`  : proofs / s with the batch resident in HBM, CUDA events on the launch stream, max over ranks.
`e2e`    : the same metric through the host-buffer C-ABI entry (mptv_verify_batch) from pinned host
           memory -- H2D of all inputs and D2H of all results inside the timed region.
`roofline`: the Keccak kernel against the measured integer-issue peak (LOP3/SHF probe run in this
           process) and against the measured HBM bandwidth (MEASURED_PEAKS.json).
`cpu_baseline`: the C restatement of the reference's CPU path (oracle/, mirroring its redundant
           hashing) on all host cores, rank 0 at N = 1 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mpt_proofs_verified_per_sec"
UNIT = "proofs/s"
I_PERM = 4320  # 32-bit integer instructions per Keccak-f


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--accounts", type=int, default=10_000_000)
    ap.add_argument("--proofs", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="proofs in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--lanes", type=int, default=0)
    return ap.parse_args()


def workload_name(a):
    return (f"config2: {a.proofs} account proofs vs synthetic {a.accounts}-account state trie "
            f"(key=keccak(address), value=account RLP), seed 2")


def build_batch(a, rank, pinned):
    from workload import gen
    t0 = time.time()
    trie = gen.SynthTrie(a.accounts, 2, kind=0)
    t1 = time.time()
    batch = gen.account_batch(trie, a.proofs, seed=2 + 1000 * rank, pinned=pinned)
    t2 = time.time()
    trie.close()
    return batch, dict(trie_build_s=round(t1 - t0, 2), proofs_emit_s=round(t2 - t1, 2))


def batch_dict(b):
    d = dict(node_bytes=b.node_bytes, node_off=b.node_off, node_len=b.node_len, proof_first=b.proof_first,
             roots=b.roots, key_bytes=b.key_bytes, key_off=b.key_off)
    if b.root_from_proof is not None:
        d["root_from_proof"] = b.root_from_proof
    return d


def sub_batch(b, n):
    """first n proofs of a batch (views, no copies)"""
    import zk_state_proofs_b200 as z
    nn = int(b.proof_first[n])
    return z.Batch(b.node_bytes, b.node_off[:nn], b.node_len[:nn], b.proof_first[:n + 1], b.roots[:32 * n],
                   b.key_bytes, b.key_off[:n + 1], None if b.root_from_proof is None else b.root_from_proof[:n], None)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t_begin, t_end):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for ts, ln in self.lines:
            if ts < t_begin or ts > t_end + 0.2:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                    power_w_max=max(power))


def cpu_baseline(b, n_sample, cores):
    """C restatement of the reference's CPU path, doing the reference's redundant work (mirror)."""
    from oracle.pyoracle import Oracle
    o = Oracle()
    s = sub_batch(b, n_sample)
    d = batch_dict(s)
    o.verify_batch(d, nthreads=cores, mirror=True)  # warm-up (page-in)
    t0 = time.perf_counter()
    st, voff, vlen, pa, pd = o.verify_batch(d, nthreads=cores, mirror=True)
    dt = time.perf_counter() - t0
    return dict(value=n_sample / dt, unit=UNIT, cores=cores, kind="port",
                sample=f"first {n_sample} proofs of the same batch, 1 timed pass after 1 warm-up pass",
                keccak_f_per_sec_algorithmic=pa / dt, keccak_f_per_sec_executed=pd / dt, seconds=round(dt, 3)), st, voff, vlen


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    b, gen_info = build_batch(a, 0, pinned=False)
    n_sample = a.cpu_sample or min(a.proofs, 100_000 * max(1, cores // 4))
    from oracle.pyoracle import Oracle
    o = Oracle()
    d = batch_dict(sub_batch(b, n_sample))
    for _ in range(a.warmup):
        o.verify_batch(d, nthreads=cores, mirror=True)
    t0 = time.perf_counter()
    pa = 0
    for _ in range(a.steps):
        _, _, _, pa, _ = o.verify_batch(d, nthreads=cores, mirror=True)
    dt = (time.perf_counter() - t0) / a.steps
    v = n_sample / dt
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=dt * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u32",
                data="synthetic", impl="reference",
                config=dict(workload=workload_name(a), reference_arm="C restatement of crypto_ops::verify_merkle_proof "
                            "(oracle/mpt_oracle.c, mirror mode: same redundant hashing as the Rust code); the Rust "
                            "reference cannot be built here (no rustc)", sample_proofs=n_sample),
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind="port",
                                  sample=f"first {n_sample} proofs of the config-2 batch per step"),
                keccak_f_per_sec=pa / dt,
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))
    return 0


def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference(a)

    import numpy as np
    import torch
    import torch.distributed as dist
    import zk_state_proofs_b200 as z

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the verifier has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ver = z.Verifier([local])  # one process per GPU; the context owns this rank's device only
    if a.lanes:
        ver.set_option("lanes_per_proof", a.lanes)
    b, gen_info = build_batch(a, rank, pinned=True)
    n_proofs, n_nodes, n_perm = b.n_proofs, b.n_nodes, b.n_perm()
    node_bytes_total = int(b.node_len.astype(np.int64).sum())
    dev = torch.device("cuda", local)

    # ---- device-resident copies (plumbing only: torch owns the buffers, libmptv.so does the work)
    def to_dev(x):
        return torch.from_numpy(x.view(np.uint8) if x.dtype != np.uint8 else x).to(dev)
    d_in = {k: to_dev(getattr(b, k)) for k in ["node_bytes", "node_off", "node_len", "proof_first", "roots",
                                               "key_bytes", "key_off"]}
    d_status = torch.zeros(n_proofs, dtype=torch.uint8, device=dev)
    d_voff = torch.zeros(n_proofs, dtype=torch.int64, device=dev)
    d_vlen = torch.zeros(n_proofs, dtype=torch.int32, device=dev)
    ptrs = {k: v.data_ptr() for k, v in d_in.items()}
    outp = dict(status=d_status.data_ptr(), value_off=d_voff.data_ptr(), value_len=d_vlen.data_ptr())
    stream = torch.cuda.Stream(device=dev)  # a real (non-default) stream: its handle is passed to the C ABI
    torch.cuda.synchronize()

    assert stream.cuda_stream != 0

    def step():
        ver.verify_batch_device(0, ptrs, n_nodes, n_proofs, outp, stream=stream.cuda_stream,
                                node_bytes_len=len(b.node_bytes))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    # ---- timed region: exactly K steps
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    ev0.record(stream)
    kernel_ms = dict(bin=0.0, keccak=0.0, parse=0.0, walk=0.0)
    launches = 0
    for _ in range(a.steps):
        step()
    ev1.record(stream)
    barrier()
    t_end = time.time()
    total_ms = ev0.elapsed_time(ev1)
    tm = ver.last_timings(0)  # per-kernel split of the LAST timed step (events on the same stream)
    launches = (tm.keccak_launches + tm.other_launches) * a.steps
    # extra (untimed) steps to average the per-kernel split
    ks = []
    for _ in range(3):
        step()
        t = ver.last_timings(0)
        ks.append((t.bin_ms, t.keccak_ms, t.parse_ms, t.walk_ms, t.total_ms))
    ks = np.array([[tm.bin_ms, tm.keccak_ms, tm.parse_ms, tm.walk_ms, tm.total_ms]] + ks)
    kavg = ks.mean(axis=0)
    clocks = sampler.stop(t_begin, t_end)

    ms_per_step = total_ms / a.steps
    if world > 1:
        t = torch.tensor([ms_per_step], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item())
        tot = torch.tensor([float(n_proofs), float(n_perm)], device=dev, dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        all_proofs, all_perm = float(tot[0].item()), float(tot[1].item())
    else:
        all_proofs, all_perm = float(n_proofs), float(n_perm)
    value = all_proofs / (ms_per_step * 1e-3)

    # ---- results of the device path, to check below
    st = d_status.cpu().numpy()
    voff = d_voff.cpu().numpy().view(np.uint64)
    vlen = d_vlen.cpu().numpy().view(np.uint32)

    # ---- e2e through the host-buffer C-ABI entry (pinned host memory, H2D + D2H inside)
    e2e = None
    if not a.no_e2e:
        for _ in range(2):
            ver.verify_batch(b)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            est, evoff, evlen = ver.verify_batch(b)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / a.steps
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = sum(int(getattr(b, k).nbytes) for k in ["node_bytes", "node_off", "node_len", "proof_first", "roots",
                                                       "key_off"]) + int(b.key_off[-1])
        e2e = dict(value=all_proofs / dt, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=13 * n_proofs,
                   ms_per_step=dt * 1e3, host_memory="pinned", timer="host wall clock around the blocking C-ABI call")
        assert (est == st).all() and (evoff == voff).all() and (evlen == vlen).all(), "host and device entries disagree"

    # ---- roofline of the dominant kernel (K1), rank 0's device
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    int_peak = max(ver.int_issue_peak(0, m) for m in (0, 2))  # lane-ops / s, measured now on this GPU
    keccak_s = float(kavg[1]) * 1e-3
    ach_int = n_perm * I_PERM * 1.0 / keccak_s
    ach_hbm = (node_bytes_total + 32 * n_nodes) / keccak_s / 1e9
    roofline = dict(
        kernel="k_keccak256_nodes", bound="int32_issue",
        achieved=ach_int / 1e12, peak=int_peak / 1e12, unit="Tlaneop/s", frac=ach_int / int_peak,
        peak_source="LOP3/SHF probe (mptv_int_issue_peak) run in this process on this GPU",
        algorithmic_ops_per_launch=n_perm * I_PERM, launch_ms=float(kavg[1]),
        keccak_f_per_sec=n_perm / keccak_s, keccak_f_per_sec_at_peak=int_peak / I_PERM,
        hbm=dict(bound="hbm", achieved=ach_hbm, peak=hbm_peak, unit="GB/s", frac=ach_hbm / hbm_peak,
                 peak_source="MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
                 algorithmic_bytes_per_launch=node_bytes_total + 32 * n_nodes),
        traffic=None,
    )

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=max(a.warmup, 3),
                ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u32",
                data="synthetic",
                config=dict(workload=workload_name(a), proofs_per_gpu=n_proofs, nodes_per_gpu=n_nodes,
                            keccak_f_per_gpu=n_perm, node_bytes_per_gpu=node_bytes_total,
                            l2="inputs (3 GB arena per step) exceed the 126 MB L2; no flush needed",
                            parallelism=f"{world} independent proof slices, one process per GPU, no collective",
                            **gen_info),
                keccak_f_per_sec=all_perm / (ms_per_step * 1e-3),
                kernel_ms=dict(bin=float(kavg[0]), keccak=float(kavg[1]), parse=float(kavg[2]), walk=float(kavg[3]),
                               total=float(kavg[4])),
                roofline=roofline, e2e=e2e, gpu_launches=int(launches), clocks=clocks)

    # ---- parity + CPU baseline (rank 0, N = 1): the oracle is the checker, never the thing measured
    if rank == 0:
        n_ok = int((st == 0).sum())
        line["verdicts"] = {z.STATUS_NAMES[i]: int(c) for i, c in enumerate(np.bincount(st, minlength=8)) if c}
        if world == 1 and not a.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_sample = a.cpu_sample or min(n_proofs, 100_000 * max(1, cores // 4))
            cb, ost, ovoff, ovlen = cpu_baseline(b, n_sample, cores)
            same = bool((ost == st[:n_sample]).all() and (ovoff == voff[:n_sample]).all() and
                        (ovlen == vlen[:n_sample]).all())
            cb["gpu_results_identical_on_sample"] = same
            line["cpu_baseline"] = cb
            if not same:
                line["parity_error"] = "GPU results differ from the oracle on the CPU-baseline sample"
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
