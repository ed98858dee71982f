#!/usr/bin/env python
"""bench.py -- MPT proofs verified / s (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path on the host cores

Default workload at every N (weak scaling, one process per GPU, no data-path collective):
BASELINE.json configs[1] -- a batch of 1 M account proofs against a synthetic 10 M-account state trie
(SURVEY.md section 8d config 2), built by workload/ (no RPC, no network).  A "step" is one pass of
the whole verification hot path (K0 binning, K1 Keccak-256 of every node, K2a decode, K2b walk)
over that batch.  The 3 GB node arena is far larger than L2 (126 MB), so every step streams from HBM.
--workload selects the other BASELINE.json configs (same JSON shape):
  config1  ONE tx-inclusion proof (index 15 of a synthetic 200-tx block trie) per call: call latency
  config3  4 M nested account + storage-slot proofs (1 M groups of 1 + 3), 80/10/10 incl/excl/mutated
  config4  tx + receipt trie root rebuild for 10 k blocks x 300 (metric: tries/s)
  config5  64 M mixed account/storage proofs, strong scaling: each rank verifies 64 M / N proofs per
           step as repeated passes over its own resident 4 M-proof pool (sampling with replacement)

`value`  : proofs / s with the batch resident in HBM, CUDA events on the launch stream, max over ranks.
`e2e`    : the same metric from the reference's own input format -- borsh(MerkleProofInput) blobs in host memory
           (crypto-ops/src/types.rs:4-9) -- through the C-ABI entry mptv_verify_borsh: flattening, H2D of everything
           the device needs and D2H of all results inside the timed region.  `e2e.roofline` puts the bytes moved
           beside the host-memory and PCIe ceilings measured in this process.
`e2e_csr`: the same from an already flattened CSR batch in pinned memory (mptv_verify_batch; last round's `e2e`).
`roofline`: the Keccak kernel against the measured integer-issue peak (LOP3/SHF probe run in this
           process) and against the measured HBM bandwidth (MEASURED_PEAKS.json).
`cpu_baseline`: the C restatement of the reference's CPU path (oracle/, mirroring its redundant
           hashing) on all host cores, rank 0 at N = 1 only.
`configs`: (default run, N = 1) compact results of BASELINE.json configs 1, 3 and 4 at full size, few steps each.
`config5`, `single_context`: (default run, N > 1) the 64 M-proof strong-scaled config over all ranks, and ONE
           context on rank 0 driving all N devices through mptv_verify_batch / mptv_verify_borsh.
"""
from __future__ import annotations

import argparse
import datetime
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
    ap.add_argument("--workload", default="config2", choices=["config1", "config2", "config3", "config4", "config5"])
    ap.add_argument("--accounts", type=int, default=10_000_000)
    ap.add_argument("--proofs", type=int, default=0, help="proofs per GPU per pass (0 = the config's size)")
    ap.add_argument("--slots", type=int, default=1_000_000, help="slots per ERC-20 storage trie (config3/5)")
    ap.add_argument("--tokens", type=int, default=64, help="distinct storage tries (config3/5)")
    ap.add_argument("--total-proofs", type=int, default=64_000_000, help="config5: proofs per step over all GPUs")
    ap.add_argument("--blocks", type=int, default=10_000, help="config4: blocks (one tx trie + one receipt trie each)")
    ap.add_argument("--per-block", type=int, default=300)
    ap.add_argument("--cpu-sample", type=int, default=0, help="proofs in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--pageable", action="store_true", help="e2e from ordinary (pageable) host memory instead of pinned")
    ap.add_argument("--borsh", action="store_true", help="(kept for old command lines: the borsh leg is always on now)")
    ap.add_argument("--shuffled", action="store_true",
                    help="config2 with the node order of EVERY proof shuffled (the reference accepts: a proof is a set): "
                         "nothing is chain-shaped, so K2f defers 100 %% of the batch to the cooperative walk K2b -- its worst case")
    ap.add_argument("--no-configs", action="store_true",
                    help="default run only: skip the compact config 1 / 3 / 4 blocks (N = 1) and config5 / single_context (N > 1)")
    ap.add_argument("--threads", type=int, default=0, help="host threads of the streamed borsh entry (0 = cores / ranks - 1)")
    ap.add_argument("--borsh-mode", type=int, default=-1, choices=[-1, 0, 1, 2],
                    help="mptv_verify_borsh mode of the e2e leg: 0 = the host flattens and aliases duplicate nodes (fewest PCIe bytes), "
                         "1 = the device flattens the page-locked blobs (the cores touch nothing), 2 = both at once; "
                         "-1 (default) = 0 on one GPU, 1 when several ranks share the host's memory system")
    ap.add_argument("--l2-fetch", type=int, default=0, help="L2 fetch granularity hint in bytes (32/64/128; 0 = leave the default)")
    ap.add_argument("--dedup", action="store_true",
                    help="SECONDARY number: hash each distinct node of the batch once (dedup_nodes option); the "
                         "headline always hashes every supplied node, as the reference does")
    return ap.parse_args()


DEFAULT_PROOFS = dict(config2=1_000_000, config3=4_000_000, config5=4_000_000)


def n_proofs_of(a):
    return a.proofs or DEFAULT_PROOFS[a.workload]


def workload_name(a):
    if a.workload == "config1":
        return ("config1: single tx inclusion proof via verify_merkle_proof (index 15 of a synthetic 200-tx block "
                "trie, txs 100-300 B, seed 1), one proof per call")
    if a.workload == "config2":
        return (f"config2: {n_proofs_of(a)} account proofs vs synthetic {a.accounts}-account state trie "
                f"(key=keccak(address), value=account RLP), seed 2" +
                (" [EVERY proof's node order shuffled: 100 % deferred to K2b, its worst case -- secondary number]" if a.shuffled else ""))
    if a.workload == "config3":
        return (f"config3: {n_proofs_of(a)} nested proofs = {n_proofs_of(a) // 4} groups x (1 account proof in a "
                f"{a.accounts}-account state trie + 3 ERC-20 slot proofs in one of {a.tokens} {a.slots}-slot storage "
                f"tries, root taken from the verified account leaf); 80% inclusion / 10% exclusion / 10% mutated "
                f"(7 mutators), seed 3; every storage key has a pre-image (an absent / wrong key is keccak of 32 random "
                f"bytes), so the batch also exists as borsh(StorageProofInput) blobs")
    if a.workload == "config5":
        return (f"config5: {a.total_proofs} mixed proofs per step over all GPUs (half account, half storage with "
                f"its account in the same shard, 2% mutated), as repeated passes over a resident "
                f"{n_proofs_of(a)}-proof pool per GPU (sampling with replacement), seed 5")
    return (f"config4: transaction + receipt trie root rebuild for {a.blocks} blocks x {a.per_block} "
            f"(tx log-normal median 180 B cap 8 KB; receipts median ~1.5 KB tail 30 KB), seed 4")


def build_batch(a, rank, pinned):
    from workload import gen
    n = n_proofs_of(a)
    t0 = time.time()
    if a.workload == "config2":
        trie = gen.SynthTrie(a.accounts, 2, kind=0)
        t1 = time.time()
        batch = gen.account_batch(trie, n, seed=2 + 1000 * rank, pinned=pinned, force_mut=6 if a.shuffled else 0)
        trie.close()
    else:
        seed = 3 if a.workload == "config3" else 5
        state, tokens = gen.make_state_and_tokens(a.accounts, a.tokens, a.slots, seed=seed)
        t1 = time.time()
        if a.workload == "config3":
            batch = gen.nested_batch(state, tokens, n // 4, seed=seed + 1000 * rank, pinned=pinned, raw_keys=True)
        else:
            batch = gen.mixed_batch(state, tokens, n, seed=seed + 1000 * rank, pinned=pinned)
        state.close()
        for t in tokens:
            t.close()
    t2 = time.time()
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
    passes = 0
    t0 = time.perf_counter()
    while True:
        st, voff, vlen, pa, pd = o.verify_batch(d, nthreads=cores, mirror=True)
        passes += 1
        dt = time.perf_counter() - t0
        if dt * cores >= 12.0 or passes >= 8:  # about 10-30 s of CPU work
            break
    dt /= passes
    return dict(value=n_sample / dt, unit=UNIT, cores=cores, kind="port",
                sample=f"first {n_sample} proofs of the same batch, {passes} timed pass(es) after 1 warm-up pass",
                keccak_f_per_sec_algorithmic=pa / dt, keccak_f_per_sec_executed=pd / dt,
                seconds=round(dt * passes, 3)), st, voff, vlen


def sample_size(a, n_proofs, cores):
    n = a.cpu_sample or min(n_proofs, 62_500 * cores)
    if a.workload in ("config3", "config5"):
        g = 4 if a.workload == "config3" else 2
        n -= n % g  # whole dependency groups
    return max(n, 1)


def rebuild_kv(a, rank, pinned):
    """config 4: per block one transaction trie and one receipt trie (interleaved), one KvBatch"""
    from workload import gen
    t0 = time.time()
    kv = gen.block_tries(a.blocks, a.per_block, "both", seed=4 + 1000 * rank, pinned=pinned)
    return kv, dict(gen_s=round(time.time() - t0, 2))


def sub_kv(kv, n_tries):
    import zk_state_proofs_b200 as z
    ni = int(kv.trie_first[n_tries])
    return z.KvBatch(kv.key_bytes, kv.key_off[:ni + 1], kv.value_bytes, kv.value_off[:ni], kv.value_len[:ni],
                     kv.trie_first[:n_tries + 1])


def config1_input():
    """200-tx block, prove index 15 (trie-utils/tests/transaction.rs:13); the proof comes from the committed
    golden vectors (built by oracle/gen_golden.py and judged by the reference ELF)"""
    import gzip
    with gzip.open(os.path.join(ROOT, "tests", "golden", "verify_vectors.json.gz"), "rb") as f:
        vs = json.loads(f.read())["vectors"]
    v = next(v for v in vs if v["tag"] == "config1/tx15")
    return bytes.fromhex(v["root"]), [bytes.fromhex(n) for n in v["proof"]], bytes.fromhex(v["key"]), bytes.fromhex(v["value"])


def run_single(a, env):
    """config 1: latency of ONE verify_merkle_proof call through the public API (host buffers in, value out)"""
    import numpy as np
    import torch
    import zk_state_proofs_b200 as z
    rank, world, local = env
    ver = z.Verifier([local])
    root, proof, key, want = config1_input()
    inp = z.MerkleProofInput(proof, root, key)
    b = z.flatten([inp])
    for _ in range(max(a.warmup, 3) * 20):
        ver.verify_batch(b)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    keep_busy(lambda: ver.verify_batch(b))
    n_calls = max(a.steps, 1) * 200
    ver.host_stats(reset=True)
    t_begin = time.time()
    t0 = time.perf_counter()
    for _ in range(n_calls):
        st, voff, vlen = ver.verify_batch(b)
    dt = (time.perf_counter() - t0) / n_calls
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end)
    tm = int(ver.host_stats(reset=True).launches // n_calls)
    assert st[0] == 0 and b.value(int(voff[0]), int(vlen[0])) == want
    # the same call without the Python mirror's per-call allocations: prebuilt ctypes structures, raw C-ABI entry
    import ctypes
    from zk_state_proofs_b200 import crypto_ops as co
    cb = co._CBatch(b.node_bytes.ctypes.data, len(b.node_bytes), b.node_off.ctypes.data, b.node_len.ctypes.data, b.n_nodes,
                    b.proof_first.ctypes.data, 1, b.roots.ctypes.data, b.key_bytes.ctypes.data, b.key_off.ctypes.data, None)
    cr = co._CResult(st.ctypes.data, voff.ctypes.data, vlen.ctypes.data)
    t0 = time.perf_counter()
    for _ in range(n_calls):
        ver.lib.mptv_verify_batch(ver.ctx, ctypes.byref(cb), ctypes.byref(cr))
    raw_dt = (time.perf_counter() - t0) / n_calls
    assert ver.verify_merkle_proof(root, proof, key) == want
    # the same proof as the bytes a guest receives (borsh(MerkleProofInput)): flattened on the calling thread, one launch
    blob = np.frombuffer(inp.to_borsh() + b"\0" * 16, np.uint8)
    boff = np.array([0, len(blob) - 16], np.uint64)
    for _ in range(50):
        ver.lib.mptv_verify_borsh(ver.ctx, blob.ctypes.data, boff.ctypes.data, 1, 0, ctypes.byref(cr))
    t0 = time.perf_counter()
    for _ in range(n_calls):
        ver.lib.mptv_verify_borsh(ver.ctx, blob.ctypes.data, boff.ctypes.data, 1, 0, ctypes.byref(cr))
    borsh_dt = (time.perf_counter() - t0) / n_calls
    assert st[0] == 0 and blob[int(voff[0]):int(voff[0]) + int(vlen[0])].tobytes() == want
    h2d = sum(int(getattr(b, k).nbytes) for k in ["node_bytes", "node_off", "node_len", "proof_first", "roots", "key_off"]) + len(key)
    line = dict(metric=METRIC, value=1.0 / dt, unit=UNIT, n_gpus=world, steps=a.steps, warmup=max(a.warmup, 3),
                ms_per_step=dt * 1e3 * 200, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u32",
                data="synthetic",
                config=dict(workload=workload_name(a), calls_per_step=200, nodes=b.n_nodes, keccak_f=b.n_perm(),
                            note=f"latency-bound: one blocking C-ABI call = {tm} kernel launch(es), inputs and the 13-byte result "
                                 "through mapped page-locked memory (no copy calls); the floor is one thread's Keccak chain "
                                 "(4320 alu instructions x 2 issue cycles = 4.4 us per rate block, 4 blocks for the longest node); "
                                 "there is no device-resident variant of a single-proof call, so value == e2e"),
                latency_us=dt * 1e6, latency_us_c_abi=raw_dt * 1e6, latency_us_borsh_c_abi=borsh_dt * 1e6, keccak_f_per_sec=b.n_perm() / dt,
                roofline=None,
                e2e=dict(value=1.0 / dt, unit=UNIT, h2d_bytes_per_step=h2d * 200, d2h_bytes_per_step=13 * 200,
                         ms_per_step=dt * 1e3 * 200, host_memory="pageable",
                         timer="host wall clock around the blocking C-ABI call"),
                gpu_launches=tm * n_calls, clocks=clocks)
    if rank == 0 and not a.no_cpu_baseline:
        from oracle.pyoracle import Oracle
        o = Oracle()
        t0 = time.perf_counter()
        reps = 20000
        for _ in range(reps):
            r = o.verify(root, proof, key, mirror=True)
        cdt = (time.perf_counter() - t0) / reps
        line["cpu_baseline"] = dict(value=1.0 / cdt, unit=UNIT, cores=1, kind="port", latency_us=cdt * 1e6,
                                    sample=f"{reps} calls of the C restatement (mirror mode) on the same proof, incl. ctypes overhead",
                                    gpu_results_identical_on_sample=bool(r[0] == 0 and r[1] == want))
    ver.close()
    return line


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    from oracle.pyoracle import Oracle
    o = Oracle()
    if a.workload == "config1":
        root, proof, key, want = config1_input()
        for _ in range(a.warmup * 200):
            o.verify(root, proof, key, mirror=True)
        n_calls = max(a.steps, 1) * 2000
        t0 = time.perf_counter()
        for _ in range(n_calls):
            r = o.verify(root, proof, key, mirror=True)
        dt = (time.perf_counter() - t0) / n_calls
        assert r[0] == 0 and r[1] == want
        emit(dict(metric=METRIC, value=1.0 / dt, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                  ms_per_step=dt * 1e3 * 2000, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u32",
                  data="synthetic", impl="reference", latency_us=dt * 1e6,
                  config=dict(workload=workload_name(a), calls_per_step=2000,
                              reference_arm="C restatement of crypto_ops::verify_merkle_proof (mirror mode), one thread, "
                                            "called through ctypes; the Rust reference cannot be built here (no rustc)"),
                  cpu_baseline=dict(value=1.0 / dt, unit=UNIT, cores=1, kind="port", sample=f"{n_calls} calls on the same proof"),
                  e2e=dict(value=1.0 / dt, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0))
        return 0
    ref_note = ("C restatement of the reference's CPU path (oracle/, mirror mode: same redundant hashing as the "
                "Rust code) on all host cores; the Rust reference cannot be built here (no rustc)")
    if a.workload == "config4":
        kv, _ = rebuild_kv(a, 0, pinned=False)
        n_sample = a.cpu_sample or min(kv.n_tries, 250 * cores)
        d = sub_kv(kv, n_sample).as_dict()
        run = lambda: o.trie_roots(d, nthreads=cores)[1]
        metric, unit, what = "mpt_trie_roots_rebuilt_per_sec", "tries/s", f"first {n_sample} tries of the config-4 batch per step"
        scaling = "weak"
    else:
        b, _ = build_batch(a, 0, pinned=False)
        n_sample = sample_size(a, b.n_proofs, cores)
        d = batch_dict(sub_batch(b, n_sample))
        run = lambda: o.verify_batch(d, nthreads=cores, mirror=True)[3]
        metric, unit, what = METRIC, UNIT, f"first {n_sample} proofs of the {a.workload} batch per step"
        scaling = "strong" if a.workload == "config5" else "weak"
    for _ in range(a.warmup):
        run()
    t0 = time.perf_counter()
    pa = 0
    for _ in range(a.steps):
        pa = run()
    dt = (time.perf_counter() - t0) / a.steps
    v = n_sample / dt
    line = dict(metric=metric, value=v, unit=unit, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=dt * 1e3, higher_is_better=True, scaling=scaling, vs_baseline=None, dtype="u32",
                data="synthetic", impl="reference",
                config=dict(workload=workload_name(a), reference_arm=ref_note, sample=n_sample),
                cpu_baseline=dict(value=v, unit=unit, cores=cores, kind="port", sample=what),
                keccak_f_per_sec=pa / dt,
                e2e=dict(value=v, unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)
    return 0


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def keccak_roofline(ver, n_perm, keccak_ms, alg_bytes, launches_note, traffic_key=None):
    peaks = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    probe = {name: ver.int_issue_peak(0, m) for m, name in ((0, "lop3"), (1, "shf"), (2, "keccak_mix_122_58"))}
    int_peak = max(probe["lop3"], probe["keccak_mix_122_58"])  # lane-ops / s, measured now on this GPU
    keccak_s = keccak_ms * 1e-3
    ach_int = n_perm * I_PERM * 1.0 / keccak_s
    ach_hbm = alg_bytes / keccak_s / 1e9
    return dict(
        kernel="k_keccak256_nodes", bound="int32_issue",
        achieved=ach_int / 1e12, peak=int_peak / 1e12, unit="Tlaneop/s", frac=ach_int / int_peak,
        peak_source="LOP3/SHF probe (mptv_int_issue_peak) run in this process on this GPU: the larger of the LOP3 and the "
                    "Keccak-mix mode; 148 SMs x 64 lanes/clk x SM clock gives the same figure",
        peak_probe_tlaneops={k: v / 1e12 for k, v in probe.items()},
        algorithmic_ops_per_launch=n_perm * I_PERM, launch_ms=keccak_ms, launches=launches_note,
        keccak_f_per_sec=n_perm / keccak_s, keccak_f_per_sec_at_peak=int_peak / I_PERM,
        hbm=dict(bound="hbm", achieved=ach_hbm, peak=hbm_peak, unit="GB/s", frac=ach_hbm / hbm_peak,
                 peak_source="MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
                 algorithmic_bytes_per_launch=alg_bytes),
        # measured DRAM bytes of ONE launch of this kernel on exactly this workload (ncu --set full), else None
        # NOT measured in this run: read from profiles/traffic.json, which holds the ncu capture named in its `source`
        traffic=dict(TRAFFIC_NOTE[traffic_key], measured_in_this_run=False) if traffic_key in TRAFFIC_NOTE else None,
    )


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of
# this kernel on this workload (profiles/); None until a capture of the current kernel exists
TRAFFIC_NOTE = {}
try:
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as _f:
        TRAFFIC_NOTE = json.load(_f)
except Exception:
    pass


def keep_busy(fn, seconds=0.25):
    """untimed extra warm-up while the clock sampler starts: the GPU must not sit idle right before the timed region
    (an idle quarter of a second lets the clocks drop, and the first timed steps then run on the ramp)"""
    import torch
    t0 = time.time()
    while time.time() - t0 < seconds:
        fn()
        torch.cuda.synchronize()


def dist_setup():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the verifier has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def reduce_max(x, world, dev):
    import torch
    import torch.distributed as dist
    if world == 1:
        return float(x)
    t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(xs, world, dev):
    import torch
    import torch.distributed as dist
    if world == 1:
        return [float(x) for x in xs]
    t = torch.tensor([float(x) for x in xs], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]


def run_rebuild(a, env):
    import numpy as np
    import torch
    import torch.distributed as dist
    import zk_state_proofs_b200 as z
    rank, world, local = env
    dev = torch.device("cuda", local)
    ver = z.Verifier([local])
    if a.l2_fetch:
        ver.set_option("l2_fetch_granularity", a.l2_fetch)
    kv, gen_info = rebuild_kv(a, rank, pinned=True)
    value_total = int(kv.value_len.astype(np.int64).sum())

    def to_dev(x):
        return torch.from_numpy(x.view(np.uint8) if x.dtype != np.uint8 else x).to(dev)
    d_in = {k: to_dev(getattr(kv, k)) for k in ["key_bytes", "key_off", "value_bytes", "value_off", "value_len", "trie_first"]}
    d_roots = torch.zeros(32 * kv.n_tries, dtype=torch.uint8, device=dev)
    ptrs = {k: v.data_ptr() for k, v in d_in.items()}
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()

    def step():
        ver.trie_roots_device(0, ptrs, kv.n_items, kv.n_tries, d_roots.data_ptr(), stream=stream.cuda_stream,
                              value_bytes_len=len(kv.value_bytes))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    keep_busy(step)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    ev0.record(stream)
    ks = []
    for _ in range(a.steps):
        step()
    ev1.record(stream)
    barrier()
    t_end = time.time()
    total_ms = ev0.elapsed_time(ev1)
    tm = ver.last_rebuild_timings(0)
    launches = (tm.keccak_launches + tm.other_launches) * a.steps
    for _ in range(3):
        step()
        t = ver.last_rebuild_timings(0)
        ks.append((t.structure_ms, t.encode_ms, t.keccak_ms, t.total_ms))
    kavg = np.array([[tm.structure_ms, tm.encode_ms, tm.keccak_ms, tm.total_ms]] + ks).mean(axis=0)
    clocks = sampler.stop(t_begin, t_end)
    ms_per_step = reduce_max(total_ms / a.steps, world, dev)
    all_tries, all_perm = reduce_sum([kv.n_tries, tm.n_perm], world, dev)
    value = all_tries / (ms_per_step * 1e-3)
    roots = d_roots.cpu().numpy().reshape(-1, 32)

    e2e = None
    if not a.no_e2e:
        ver.trie_roots(kv)
        barrier()
        t0 = time.perf_counter()
        for _ in range(min(a.steps, 3)):
            eroots = ver.trie_roots(kv)
        torch.cuda.synchronize()
        dt = reduce_max((time.perf_counter() - t0) / min(a.steps, 3), world, dev)
        h2d = sum(int(getattr(kv, k).nbytes) for k in ["key_off", "value_bytes", "value_off", "value_len", "trie_first"]) + int(kv.key_off[-1])
        e2e = dict(value=all_tries / dt, unit="tries/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=32 * kv.n_tries,
                   ms_per_step=dt * 1e3, host_memory="pinned values, pageable index arrays",
                   timer="host wall clock around the blocking C-ABI call")
        assert (eroots == roots).all(), "host and device entries disagree"

    roofline = keccak_roofline(ver, tm.n_perm, float(kavg[2]), int(tm.arena_bytes) + 32 * int(tm.n_hashed),
                               f"{tm.keccak_launches} (one per trie level), summed")
    line = dict(metric="mpt_trie_roots_rebuilt_per_sec", value=value, unit="tries/s", n_gpus=world, steps=a.steps,
                warmup=max(a.warmup, 3), ms_per_step=ms_per_step, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="u32", data="synthetic",
                config=dict(workload=workload_name(a), tries_per_gpu=kv.n_tries, items_per_gpu=kv.n_items,
                            value_bytes_per_gpu=value_total, trie_nodes_per_gpu=int(tm.n_nodes),
                            hashed_nodes_per_gpu=int(tm.n_hashed), keccak_f_per_gpu=int(tm.n_perm), levels=int(tm.levels),
                            l2="inputs (GBs of values per step) exceed the 126 MB L2; no flush needed",
                            parallelism=f"{world} independent trie slices, one process per GPU, no collective", **gen_info),
                keccak_f_per_sec=all_perm / (ms_per_step * 1e-3),
                leaves_per_sec=all_tries * a.per_block / (ms_per_step * 1e-3),
                kernel_ms=dict(structure=float(kavg[0]), encode=float(kavg[1]), keccak=float(kavg[2]), total=float(kavg[3])),
                roofline=roofline, e2e=e2e, gpu_launches=int(launches), clocks=clocks)
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            from oracle.pyoracle import Oracle
            o = Oracle()
            cores = os.cpu_count() or 1
            n_sample = a.cpu_sample or min(kv.n_tries, 250 * cores)
            d = sub_kv(kv, n_sample).as_dict()
            t0 = time.perf_counter()
            oroots, pa, nh = o.trie_roots(d, nthreads=cores)
            dt = time.perf_counter() - t0
            same = bool((oroots == roots[:n_sample]).all())
            line["cpu_baseline"] = dict(value=n_sample / dt, unit="tries/s", cores=cores, kind="port",
                                        sample=f"first {n_sample} tries of the same batch, 1 timed pass",
                                        keccak_f_per_sec=pa / dt, seconds=round(dt, 3),
                                        gpu_results_identical_on_sample=same)
            if not same:
                line["parity_error"] = "GPU roots differ from the oracle on the CPU-baseline sample"
    del d_in, d_roots
    ver.close()
    torch.cuda.empty_cache()
    return line


def emit(line):
    """the ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) to stderr"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def flatten_threads(a, world):
    """host threads for the streamed borsh entry: this rank's share of the cores, less two when there are at least
    eight (the submitter thread and this process's other threads; the library's own default), else less one"""
    if a.threads:
        return a.threads
    cores = os.cpu_count() or 1
    share = max(1, cores // max(world, 1))
    if share >= 8:
        return share - 2
    return max(1, share - 1) if share > 2 else share


def h2d_peak_gbs(dev, world):
    """pinned host -> device copy bandwidth measured now, on every rank at the same time (what N concurrent feeds get)"""
    import torch
    import torch.distributed as dist
    n = 1 << 29
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 4
    del h, d
    return n / dt / 1e9


def e2e_from_borsh(a, ver, b, env, dev, st, voff, vlen, steps):
    """blobs in host memory -> verdicts: the streamed C-ABI entry, with the bytes it moved and the ceilings beside it"""
    import numpy as np
    import torch
    import torch.distributed as dist
    import zk_state_proofs_b200 as z
    from workload import gen
    rank, world, local = env
    blobs, boff = gen.batch_to_borsh(b, pinned=not a.pageable)  # page-locked unless --pageable (the device-flatten mode needs it)
    th = flatten_threads(a, world)
    mode = a.borsh_mode if a.borsh_mode >= 0 else (1 if world > 1 and not a.pageable else 0)
    ver.set_option("borsh_mode", mode)
    for _ in range(2):
        bst, bvoff, bvlen = ver.verify_borsh(blobs, boff, threads=th)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ver.host_stats(reset=True)
    t0 = time.perf_counter()
    marks = [t0]
    for _ in range(steps):
        bst, bvoff, bvlen = ver.verify_borsh(blobs, boff, threads=th)
        marks.append(time.perf_counter())
    dt = reduce_max((marks[-1] - t0) / steps, world, dev)
    hs = ver.host_stats(reset=True)
    assert (bst == st).all() and (bvlen == vlen).all(), "borsh stream and device entry disagree"
    for i in np.nonzero(bst == 0)[0][:2000]:
        assert blobs[int(bvoff[i]):int(bvoff[i]) + int(bvlen[i])].tobytes() == b.value(int(voff[i]), int(vlen[i]))
    # ceilings, measured now on every rank at once
    read_gbs, copy_gbs = z.host_bw_probe(th, 128 << 20)
    h2d_gbs = h2d_peak_gbs(dev, world)
    flat_s = min(z.borsh_flatten_probe(blobs, boff, threads=th, chunk_bytes=32 << 20)[0] for _ in range(3))
    ver.set_option("borsh_mode", -1)
    all_proofs, blob_bytes, h2d, d2h, read_sum, copy_sum, h2d_sum, launches = reduce_sum(
        [b.n_proofs, len(blobs), hs.h2d_bytes / steps, hs.d2h_bytes / steps, read_gbs, copy_gbs, h2d_gbs, hs.launches / steps], world, dev)
    placed = hs.node_bytes_placed / steps
    # every byte the host memory system moves per step on this rank: the blobs are read once by the cores; what is
    # staged is written once by the cores and read once by the DMA engine; in pull mode the placed node bytes are not
    # staged but read once more, by the device, straight from the blobs
    pulled = hs.node_bytes_placed / steps if hs.pull_chunks else 0
    if hs.device_chunks == hs.chunks:   # device flatten: the blobs are read once, by the DMA engine; the cores touch nothing
        dram = hs.h2d_bytes / steps
    else:
        dram = len(blobs) + 2 * (hs.h2d_bytes / steps - pulled) + pulled
    dram_all = reduce_sum([dram], world, dev)[0]
    return dict(
        value=all_proofs / dt, unit=UNIT, ms_per_step=dt * 1e3, entry="mptv_verify_borsh",
        step_ms_rank0=[round((y - x) * 1e3, 2) for x, y in zip(marks, marks[1:])],
        host_memory=("pageable blobs, page-locked staging inside the library" if a.pageable else "page-locked blobs"),
        borsh_mode={0: "0: the host flattens (one pass, byte-identical nodes of a chunk aliased), staging copied by the DMA engine",
                    1: "1: the blobs cross PCIe as they are and the device flattens them (borsh_kernels.cu): several ranks share one "
                       "host memory system, so the cores touch nothing",
                    2: "2: both pipelines at once"}[mode],
        device_chunks_per_step=int(hs.device_chunks / steps),
        pull_chunks_per_step=int(hs.pull_chunks / steps), chunks_per_step=int(hs.chunks / steps),
        input_bytes_per_step=int(blob_bytes), h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h), host_threads_per_rank=th,
        timer="host wall clock around the blocking C-ABI call, max over ranks",
        transfer_dedup=dict(nodes=int(hs.nodes / steps), nodes_aliased=int(hs.nodes_aliased / steps),
                            node_bytes_supplied=int(hs.node_bytes_supplied / steps), node_bytes_placed=int(placed),
                            note="byte-identical nodes of a chunk cross PCIe once (exact compare); every supplied node is still hashed on the device"),
        host_ms=dict(flatten=hs.flatten_us / steps / 1e3, wait=hs.wait_us / steps / 1e3, map_results=hs.map_us / steps / 1e3,
                     call=hs.call_us / steps / 1e3, host_stage_alone=flat_s * 1e3),
        roofline=dict(**(dict(bound="host_dram", achieved=dram_all / dt / 1e9, peak=2 * copy_sum, frac=dram_all / dt / 1e9 / (2 * copy_sum))
                         if mode != 1 or dram_all / (2 * copy_sum) >= h2d / h2d_sum else
                         # device flatten: every byte is one DMA read; the links bind before the memory system does
                         dict(bound="pcie", achieved=h2d / dt / 1e9, peak=h2d_sum, frac=h2d / dt / 1e9 / h2d_sum)),
                      unit="GB/s", bytes_per_step=int(dram_all), peak_source=f"host_dram: mptv_host_bw_probe run now with the same {th} threads on every rank at once: "
                      "read + non-temporal write copy, memory traffic = 2 x payload, summed over ranks; pcie: pinned-copy loop on all ranks at once",
                      host_read=dict(achieved=blob_bytes / dt / 1e9, peak=read_sum, unit="GB/s", frac=blob_bytes / dt / 1e9 / read_sum,
                                     note="blob bytes read by the cores vs the read-only probe"),
                      pcie=dict(achieved=h2d / dt / 1e9, peak=h2d_sum, unit="GB/s", frac=h2d / dt / 1e9 / h2d_sum,
                                note="H2D bytes vs pinned-copy bandwidth measured now on all ranks at once, summed"),
                      limiter=("host memory bandwidth: the cores must read every supplied byte once to compare it" if mode == 0 else
                               "the host's memory system / PCIe: every blob byte is read once by the DMA engines of all ranks")),
        gpu_launches_per_step=int(launches))


def e2e_from_storage_borsh(a, ver, b, env, dev, st, voff, vlen, steps):
    """the nested workload from ITS wire format: one borsh(StorageProofInput) blob per group in ordinary host memory ->
    the storage guest's flow for every input (mptv_verify_storage_borsh), flattening, copies and results inside"""
    import numpy as np
    import torch
    import torch.distributed as dist
    from workload import gen
    rank, world, local = env
    blobs, boff, gf = gen.batch_to_storage_borsh(b, pinned=False)
    th = flatten_threads(a, world)
    call = lambda: ver.verify_storage_borsh(blobs, boff, threads=th, n_proofs=b.n_proofs)
    for _ in range(2):
        pf, ist, bst, bvoff, bvlen = call()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ver.host_stats(reset=True)
    t0 = time.perf_counter()
    marks = [t0]
    for _ in range(steps):
        pf, ist, bst, bvoff, bvlen = call()
        marks.append(time.perf_counter())
    dt = reduce_max((marks[-1] - t0) / steps, world, dev)
    hs = ver.host_stats(reset=True)
    assert (pf == gf).all() and (bst == st).all() and (bvlen == vlen).all(), "storage stream and device entry disagree"
    for i in np.nonzero(bst == 0)[0][:2000]:
        assert blobs[int(bvoff[i]):int(bvoff[i]) + int(bvlen[i])].tobytes() == b.value(int(voff[i]), int(vlen[i]))
    # the guest's outcome per input: the first failing proof of the group
    first_bad = np.array([next((int(s) for s in st[int(gf[g]):int(gf[g + 1])] if s), 0) for g in range(0, len(gf) - 1, 997)])
    assert (ist[::997] == first_bad).all()
    all_proofs, blob_bytes, h2d, d2h, launches = reduce_sum(
        [b.n_proofs, len(blobs), hs.h2d_bytes / steps, hs.d2h_bytes / steps, hs.launches / steps], world, dev)
    return dict(value=all_proofs / dt, unit=UNIT, ms_per_step=dt * 1e3, entry="mptv_verify_storage_borsh",
                step_ms_rank0=[round((y - x) * 1e3, 2) for x, y in zip(marks, marks[1:])],
                inputs_per_step=int(len(gf) - 1), host_memory="pageable blobs, page-locked staging inside the library",
                input_bytes_per_step=int(blob_bytes), h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                host_threads_per_rank=th, chunks_per_step=int(hs.chunks / steps),
                timer="host wall clock around the blocking C-ABI call (index pass + stream), max over ranks",
                transfer_dedup=dict(nodes=int(hs.nodes / steps), nodes_aliased=int(hs.nodes_aliased / steps),
                                    node_bytes_supplied=int(hs.node_bytes_supplied / steps),
                                    node_bytes_placed=int(hs.node_bytes_placed / steps)),
                host_ms=dict(index_pass=hs.index_us / steps / 1e3, flatten=hs.flatten_us / steps / 1e3, wait=hs.wait_us / steps / 1e3,
                             map_results=hs.map_us / steps / 1e3, stream=hs.call_us / steps / 1e3),
                gpu_launches_per_step=int(launches),
                note="storage keys cross PCIe un-hashed and are hashed on the device (digest_keccak(&key), main.rs:26); "
                     "storage roots are taken on the device from the verified account leaves")


def e2e_from_csr(ver, b, names, env, dev, st, voff, vlen, steps, passes, pageable):
    import torch
    import torch.distributed as dist
    rank, world, local = env
    for _ in range(2):
        ver.verify_batch(b)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ver.host_stats(reset=True)
    t0 = time.perf_counter()
    for _ in range(steps * passes):
        est, evoff, evlen = ver.verify_batch(b)
    torch.cuda.synchronize()
    dt = reduce_max((time.perf_counter() - t0) / steps, world, dev)
    hs = ver.host_stats(reset=True)
    assert (est == st).all() and (evoff == voff).all() and (evlen == vlen).all(), "host and device entries disagree"
    all_proofs, h2d, d2h = reduce_sum([b.n_proofs * passes, hs.h2d_bytes / steps, hs.d2h_bytes / steps], world, dev)
    return dict(value=all_proofs / dt, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h), ms_per_step=dt * 1e3,
                entry="mptv_verify_batch", host_memory="pageable" if pageable else "pinned",
                timer="host wall clock around the blocking C-ABI call, max over ranks", pcie_gbs=h2d / dt / 1e9)


def run_verify(a, env, want_e2e=True, want_cpu=True):
    """configs 2, 3, 5: device-resident steps, the host-fed entries, roofline, parity sample"""
    import numpy as np
    import torch
    import torch.distributed as dist
    import zk_state_proofs_b200 as z

    rank, world, local = env
    ver = z.Verifier([local])  # one process per GPU; the context owns this rank's device only
    if a.lanes:
        ver.set_option("lanes_per_proof", a.lanes)
    if a.dedup:
        ver.set_option("dedup_nodes", 1)
    if a.l2_fetch:
        ver.set_option("l2_fetch_granularity", a.l2_fetch)
    b, gen_info = build_batch(a, rank, pinned=not a.pageable)
    n_proofs, n_nodes, n_perm = b.n_proofs, b.n_nodes, b.n_perm()
    node_bytes_total = int(b.node_len.astype(np.int64).sum())
    dev = torch.device("cuda", local)
    # config 5: a step = total_proofs / world proofs on this rank = `passes` passes over its resident pool
    passes = 1
    if a.workload == "config5":
        passes = max(1, round(a.total_proofs / world / n_proofs))

    # ---- device-resident copies (plumbing only: torch owns the buffers, libmptv.so does the work)
    def to_dev(x):
        return torch.from_numpy(x.view(np.uint8) if x.dtype != np.uint8 else x).to(dev)
    names = ["node_bytes", "node_off", "node_len", "proof_first", "roots", "key_bytes", "key_off"]
    if b.root_from_proof is not None:
        names.append("root_from_proof")
    d_in = {k: to_dev(getattr(b, k)) for k in names}
    d_status = torch.zeros(n_proofs, dtype=torch.uint8, device=dev)
    d_voff = torch.zeros(n_proofs, dtype=torch.int64, device=dev)
    d_vlen = torch.zeros(n_proofs, dtype=torch.int32, device=dev)
    ptrs = {k: v.data_ptr() for k, v in d_in.items()}
    outp = dict(status=d_status.data_ptr(), value_off=d_voff.data_ptr(), value_len=d_vlen.data_ptr())
    stream = torch.cuda.Stream(device=dev)  # a real (non-default) stream: its handle is passed to the C ABI
    torch.cuda.synchronize()

    assert stream.cuda_stream != 0

    def step():
        for _ in range(passes):
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
    keep_busy(step)
    # ---- timed region: exactly K steps
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.time()
    ev0.record(stream)
    for _ in range(a.steps):
        step()
    ev1.record(stream)
    barrier()
    t_end = time.time()
    total_ms = ev0.elapsed_time(ev1)
    tm = ver.last_timings(0)  # per-kernel split of the LAST timed pass (events on the same stream)
    launches = (tm.keccak_launches + tm.other_launches) * a.steps * passes
    # extra (untimed) passes to average the per-kernel split
    ks = []
    for _ in range(3):
        ver.verify_batch_device(0, ptrs, n_nodes, n_proofs, outp, stream=stream.cuda_stream,
                                node_bytes_len=len(b.node_bytes))
        t = ver.last_timings(0)
        ks.append((t.bin_ms, t.keccak_ms, t.parse_ms, t.walk_ms, t.total_ms))
    ks = np.array([[tm.bin_ms, tm.keccak_ms, tm.parse_ms, tm.walk_ms, tm.total_ms]] + ks)
    kavg = ks.mean(axis=0)
    clocks = sampler.stop(t_begin, t_end)

    ms_per_step = reduce_max(total_ms / a.steps, world, dev)
    all_proofs, all_perm = reduce_sum([n_proofs * passes, n_perm * passes], world, dev)
    value = all_proofs / (ms_per_step * 1e-3)

    # ---- results of the device path, to check below
    st = d_status.cpu().numpy()
    voff = d_voff.cpu().numpy().view(np.uint64)
    vlen = d_vlen.cpu().numpy().view(np.uint32)

    # ---- end to end through the host-fed C-ABI entries
    e2e = e2e_csr = None
    if want_e2e and not a.no_e2e:
        e_steps = a.steps if passes == 1 else 1
        e2e_csr = e2e_from_csr(ver, b, names, env, dev, st, voff, vlen, e_steps, passes, a.pageable)
        if passes == 1 and b.root_from_proof is None:
            # the reference's input format: borsh(MerkleProofInput) blobs (MerkleProofInput has no dependency field,
            # so the nested configs have no blob form)
            e2e = e2e_from_borsh(a, ver, b, env, dev, st, voff, vlen, a.steps)
        elif passes == 1 and getattr(b, "raw_keys", None) is not None:
            # the nested workload's own wire format: one borsh(StorageProofInput) per group (types.rs:11-19)
            e2e = e2e_from_storage_borsh(a, ver, b, env, dev, st, voff, vlen, a.steps)
        else:
            e2e = dict(e2e_csr, note="nested / multi-pass workload: StorageProofInput groups have no single-blob MerkleProofInput form, "
                                     "so the end-to-end leg runs from the flattened CSR batch in pinned host memory")

    # ---- roofline of the dominant kernel (K1), rank 0's device
    perm_executed = int(tm.n_unique_perm) if a.dedup else n_perm
    full_config2 = a.workload == "config2" and not a.dedup and n_proofs == 1_000_000 and a.accounts == 10_000_000
    roofline = keccak_roofline(ver, perm_executed, float(kavg[1]), node_bytes_total + 32 * n_nodes, "1 per pass",
                               "k_keccak256_nodes" if full_config2 else None)

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=max(a.warmup, 3),
                ms_per_step=ms_per_step, higher_is_better=True, scaling="strong" if a.workload == "config5" else "weak",
                vs_baseline=None, dtype="u32", data="synthetic",
                config=dict(workload=workload_name(a), proofs_per_gpu=n_proofs, passes_per_step=passes,
                            nodes_per_gpu=n_nodes, keccak_f_per_gpu=n_perm, node_bytes_per_gpu=node_bytes_total,
                            l2="inputs (GBs of node bytes per pass) exceed the 126 MB L2; no flush needed",
                            parallelism=f"{world} independent proof slices, one process per GPU, no collective",
                            **gen_info),
                keccak_f_per_sec=all_perm / (ms_per_step * 1e-3),
                kernel_ms=dict(bin=float(kavg[0]), keccak=float(kavg[1]), parse=float(kavg[2]), walk=float(kavg[3]),
                               total=float(kavg[4])),
                roofline=roofline, e2e=e2e, e2e_csr=e2e_csr, e2e_borsh=e2e if (e2e and e2e.get("entry") == "mptv_verify_borsh") else None,
                gpu_launches=int(launches), clocks=clocks)

    if a.dedup:
        line["dedup"] = dict(unique_nodes=int(tm.n_unique_nodes), nodes=n_nodes, keccak_f_executed=int(tm.n_unique_perm),
                             keccak_f_algorithmic=n_perm,
                             note="SECONDARY NUMBER: each distinct node of the batch hashed once and its digest shared "
                                  "(results identical). keccak_f_per_sec above still counts the algorithmic W_perm "
                                  "(every supplied node); the roofline uses the executed count. The headline run never "
                                  "deduplicates.")
        line["config"]["workload"] += " [DEDUP RUN -- secondary number]"
    # ---- parity + CPU baseline (rank 0, N = 1): the oracle is the checker, never the thing measured
    if rank == 0:
        line["verdicts"] = {z.STATUS_NAMES[i]: int(c) for i, c in enumerate(np.bincount(st, minlength=8)) if c}
        if world == 1 and want_cpu and not a.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_sample = sample_size(a, n_proofs, cores)
            cb, ost, ovoff, ovlen = cpu_baseline(b, n_sample, cores)
            same = bool((ost == st[:n_sample]).all() and (ovoff == voff[:n_sample]).all() and
                        (ovlen == vlen[:n_sample]).all())
            cb["gpu_results_identical_on_sample"] = same
            line["cpu_baseline"] = cb
            if not same:
                line["parity_error"] = "GPU results differ from the oracle on the CPU-baseline sample"
    return line, ver, b


def single_context(a, env, b):
    """N > 1, rank 0 only: ONE context that owns all N devices drives them through the host-buffer entries
    (mptv_create(ids, N) -> mptv_verify_batch / mptv_verify_borsh: per-device host threads and streams, no collective)
    while the other ranks wait at a barrier."""
    import numpy as np
    import zk_state_proofs_b200 as z
    from workload import gen
    rank, world, local = env
    ver = z.Verifier(list(range(world)))
    out = dict(devices=ver.device_count, proofs=b.n_proofs)
    ref = None
    for name in ("mptv_verify_batch", "mptv_verify_borsh"):
        if name == "mptv_verify_batch":
            call = lambda: ver.verify_batch(b)
        else:
            blobs, boff = gen.batch_to_borsh(b, pinned=not a.pageable)
            ver.set_option("borsh_mode", a.borsh_mode)  # -1: the library's own choice (device flatten for page-locked blobs on several devices)
            call = lambda: ver.verify_borsh(blobs, boff, threads=a.threads)
        for _ in range(2):
            call()
        ver.host_stats(reset=True)
        steps = 3
        t0 = time.perf_counter()
        for _ in range(steps):
            res = call()
        dt = (time.perf_counter() - t0) / steps
        hs = ver.host_stats(reset=True)
        if ref is None:
            ref = res
        same = bool((res[0] == ref[0]).all() and (res[2] == ref[2]).all())
        out[name] = dict(value=b.n_proofs / dt, unit=UNIT, ms_per_step=dt * 1e3, h2d_bytes_per_step=int(hs.h2d_bytes / steps),
                         d2h_bytes_per_step=int(hs.d2h_bytes / steps), chunks_per_step=int(hs.chunks / steps),
                         same_verdicts_as_first_entry=same)
    out["verdicts"] = {z.STATUS_NAMES[i]: int(c) for i, c in enumerate(np.bincount(ref[0], minlength=8)) if c}
    out["note"] = ("strong-scaled: the rank-0 batch of the weak-scaled line cut into N device slices by ONE context; "
                   "host wall clock around the blocking call; all N GPUs are fed from this one process")
    ver.close()
    return out


def compact(line, keys=("value", "unit", "ms_per_step", "kernel_ms", "keccak_f_per_sec", "latency_us", "latency_us_c_abi", "latency_us_borsh_c_abi", "verdicts", "gpu_launches",
                        "leaves_per_sec", "parity_error")):
    out = {k: line[k] for k in keys if k in line and line[k] is not None}
    out["workload"] = line["config"]["workload"]
    out["steps"] = line["steps"]
    if line.get("roofline"):
        out["roofline_frac"] = line["roofline"]["frac"]
    if line.get("e2e"):
        out["e2e"] = {k: line["e2e"][k] for k in ("value", "unit", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "entry")
                      if k in line["e2e"]}
    if line.get("cpu_baseline"):
        cb = line["cpu_baseline"]
        out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "latency_us", "gpu_results_identical_on_sample") if k in cb}
        out["gpu_results_identical_on_sample"] = cb.get("gpu_results_identical_on_sample")
    return out


def sub_args(a, **kw):
    s = argparse.Namespace(**vars(a))
    for k, v in kw.items():
        setattr(s, k, v)
    return s


def main():
    global _REAL_STDOUT
    a = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # libraries that print to fd 1 (e.g. "NCCL version ...") now land on stderr
    if a.impl == "reference":
        return run_reference(a)
    import torch
    import torch.distributed as dist
    env = dist_setup()
    rank, world, local = env
    if a.workload == "config4":
        line = run_rebuild(a, env)
    elif a.workload == "config1":
        line = run_single(a, env)
    else:
        line, ver, b = run_verify(a, env)
        default_run = a.workload == "config2" and not a.no_configs and not a.dedup and not a.proofs and not a.shuffled
        if default_run and world == 1:
            # the other BASELINE.json configs at full size, few steps each (their own launches of this same program:
            # `--workload configN` gives the full line)
            ver.close()
            del b
            torch.cuda.empty_cache()
            cfgs = {}
            for name, fn, kw in (("config1", run_single, dict(steps=5, warmup=3)),
                                 ("config3", None, dict(steps=3, warmup=3)),
                                 ("config4", run_rebuild, dict(steps=3, warmup=3))):
                sa = sub_args(a, workload=name, **kw)
                try:
                    if fn is None:
                        l3, v3, b3 = run_verify(sa, env)
                        v3.close()
                        del b3
                    else:
                        l3 = fn(sa, env)
                    cfgs[name] = compact(l3)
                except Exception as e:  # a failed side block must not lose the headline line
                    cfgs[name] = dict(error=f"{type(e).__name__}: {e}")
                torch.cuda.empty_cache()
            line["configs"] = cfgs
        elif default_run and world > 1:
            # (a) ONE context on rank 0 drives all N devices; the other ranks wait
            # (they wait on the rendezvous store, in a socket read: an NCCL barrier would park a spinning kernel on every
            # other GPU, and this context's kernels would have to share those GPUs with it, time slice by time slice)
            store = dist.distributed_c10d._get_default_store()
            if rank == 0:
                try:
                    line["single_context"] = single_context(a, env, b)
                except Exception as e:
                    line["single_context"] = dict(error=f"{type(e).__name__}: {e}")
                store.set("mptv_single_context_done", "1")
            else:
                store.wait(["mptv_single_context_done"], datetime.timedelta(minutes=20))
            dist.barrier()
            ver.close()
            del b
            torch.cuda.empty_cache()
            # (b) BASELINE.json config 5: 64 M mixed proofs per step, strong-scaled over the ranks
            try:
                l5, v5, b5 = run_verify(sub_args(a, workload="config5", steps=3, warmup=3), env, want_cpu=False)
                v5.close()
                if rank == 0:
                    line["config5"] = compact(l5)
                    line["config5"]["scaling"] = "strong"
            except Exception as e:
                if rank == 0:
                    line["config5"] = dict(error=f"{type(e).__name__}: {e}")
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
