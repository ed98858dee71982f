// rebuild_api.cu -- C ABI of the trie rebuild (include/mptv.h: mptv_trie_roots*): the host-side
// driver of the level-synchronous K4 pipeline in rebuild_kernels.cu.  Replaces the call sequence
// EthTrie::new / insert x n / root_hash of /root/reference/trie-utils/src/proofs/transaction.rs:41-66
// and proofs/receipt.rs:49-84 for a batch of independent tries.  No CPU fallback: every hash, every
// RLP byte and the trie structure itself are produced on the device.
#include <string.h>

#include <thread>
#include <vector>

#include "ctx.h"

using namespace mptv;

namespace {

int rebuild_on_device(mptv_ctx* ctx, Device& d, const TrieBatchDev& in, uint8_t* roots32, cudaStream_t st) {
  Rebuild& rb = d.rb;
  if (!rb.ev_begin) {
    CK(cudaEventCreate(&rb.ev_begin));
    CK(cudaEventCreate(&rb.ev_struct));
    CK(cudaEventCreate(&rb.ev_end));
  }
  const size_t N = (size_t)in.n_items, T = in.n_tries;
  CK(rb.sum.reserve(sizeof(TrieSummary)));
  CK(rb.h_sum.reserve(sizeof(TrieSummary)));
  CK(rb.rec.reserve(sizeof(uint4) * 3 * N + 16));
  CK(rb.off.reserve(8 * 3 * N + 8));
  CK(rb.len.reserve(4 * 3 * N + 4));
  CK(rb.digests.reserve(32 * 3 * N + 32));
  CK(rb.lvl_list.reserve(4 * 3 * N + 4));
  CK(rb.tcount.reserve(4 * T + 4));
  CK(rb.bins.reserve(kBinScratchWords * sizeof(uint32_t)));
  TrieWork w;
  w.rec = rb.rec.as<uint4>(); w.off = rb.off.as<uint64_t>(); w.len = rb.len.as<uint32_t>();
  w.digests = rb.digests.as<uint8_t>(); w.tcount = rb.tcount.as<uint32_t>(); w.lvl_list = rb.lvl_list.as<uint32_t>();
  w.sum = rb.sum.as<TrieSummary>();
  TrieSummary* hs = reinterpret_cast<TrieSummary*>(rb.h_sum.p);

  CK(cudaEventRecord(rb.ev_begin, st));
  // ---- pass 0: largest trie / longest key (sizes the shared-memory sort; refuses what K4 cannot hold)
  CK(launch_trie_scan_input(in, w.sum, st));
  CK(cudaMemcpyAsync(hs, w.sum, sizeof(TrieSummary), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (hs->max_items > (uint32_t)kTrieMaxItems || hs->max_key_len > (uint32_t)kTrieMaxKeyLen) {
    fail_msg(ctx, MPTV_ERR_ARG, "mptv_trie_roots: a trie has more than 8192 items or a key longer than 32 bytes");
    return MPTV_ERR_ARG;
  }
  // ---- pass 1: structure, exact node sizes, level lists
  CK(launch_trie_structure(in, w, hs->max_items, st));
  CK(cudaMemcpyAsync(hs, w.sum, sizeof(TrieSummary), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(rb.ev_struct, st));
  CK(rb.arena.reserve((size_t)hs->arena_bytes + 64));
  uint32_t levels = 0, max_level_nodes = 0;
  for (int h = 0; h < kMaxLevels; h++) {
    const uint32_t c = hs->lvl_count[2 * h] + hs->lvl_count[2 * h + 1];
    if (c) levels = h + 1;
    if (hs->lvl_count[2 * h] > max_level_nodes) max_level_nodes = hs->lvl_count[2 * h];
  }
  CK(rb.order.reserve(4 * (size_t)max_level_nodes + 4));
  while (rb.lvl_ev.size() < 3 * (size_t)levels) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    rb.lvl_ev.push_back(e);
  }
  // ---- pass 2: bottom-up, one encode launch + one hash launch per level
  uint32_t start = 0, klaunch = 0, olaunch = 4;
  for (uint32_t h = 0; h < levels; h++) {
    const uint32_t nh = hs->lvl_count[2 * h], ni = hs->lvl_count[2 * h + 1];
    const uint32_t* list = w.lvl_list + start;
    // height 0 = leaves.  Fused path: the >= 32-byte leaves are hashed straight from the value arena
    // (K1L) and never written; only the inline (< 32 byte) leaves are encoded, for their parents to embed.
    const bool fused = h == 0 && ctx->fused_leaf_hash;
    CK(cudaEventRecord(rb.lvl_ev[3 * h], st));
    if (fused) CK(launch_trie_encode(in, w, list + nh, ni, rb.arena.as<uint8_t>(), st));
    else CK(launch_trie_encode(in, w, list, nh + ni, rb.arena.as<uint8_t>(), st));
    if (fused ? ni : nh + ni) olaunch++;
    CK(cudaEventRecord(rb.lvl_ev[3 * h + 1], st));
    if (nh) {
      const uint32_t* ord = list;
      if (ctx->binning && nh >= 4096) {  // leaf levels: converge the warps on equal rate-block counts
        CK(launch_bin_nodes(w.len, list, nh, rb.bins.as<uint32_t>(), rb.order.as<uint32_t>(), st, nullptr, nullptr,
                            ctx->long_leaf_bin));
        ord = rb.order.as<uint32_t>();
        olaunch += 3;
      }
      uint32_t* tiles = rb.bins.as<uint32_t>() + 2 * kNumBins;
      // binned levels: the long nodes (> 32 rate blocks) are hashed first, in a launch of their own
      const uint32_t* split = ord != list ? bin_split_word(rb.bins.as<uint32_t>()) : nullptr;
      if (fused) CK(launch_keccak256_leaves(in, w.rec, w.len, ord, nh, w.digests, tiles, d.sm_count, st, split, ctx->long_leaf_ctas));
      else CK(launch_keccak256_nodes(rb.arena.as<uint8_t>(), 0, w.off, w.len, ord, nh, w.digests, nullptr, tiles, d.sm_count, st,
                                     split, ctx->long_leaf_ctas));
      klaunch += split ? 2 : 1;
    }
    CK(cudaEventRecord(rb.lvl_ev[3 * h + 2], st));
    start += nh + ni;
  }
  CK(launch_trie_roots(in, w, roots32, st));
  olaunch++;
  CK(cudaEventRecord(rb.ev_end, st));
  rb.leaves_in_arena = !ctx->fused_leaf_hash;
  rb.levels = levels; rb.keccak_launches = klaunch; rb.other_launches = olaunch; rb.have_timing = true;
  rb.n_nodes = hs->n_nodes; rb.n_hashed = hs->nodes_hashed; rb.n_perm = hs->perms; rb.arena_bytes = hs->arena_bytes;
  return MPTV_OK;
}

// queues the copy of tries [cs, ce) of a host batch into one of the two input stages (offsets rebased
// to the chunk through page-locked staging, value bytes straight from the caller's arena); asynchronous
// on `st`, completion = stage.up
int upload_kv(mptv_ctx* ctx, Device& d, const mptv_kv_batch* in, uint64_t cs, uint64_t ce, KvStage& g, TrieBatchDev& b,
              cudaStream_t st) {
  const uint32_t i0 = in->trie_first[cs], i1 = in->trie_first[ce];
  const uint64_t ni = i1 - i0, nt = ce - cs;
  const uint64_t b0 = i0 < in->n_items ? in->value_off[i0] : in->value_bytes_len;
  uint64_t b1 = b0;
  if (ni) b1 = in->value_off[i1 - 1] + in->value_len[i1 - 1];
  b1 = (b1 + 15) & ~15ull;
  if (b1 > ((in->value_bytes_len + 15) & ~15ull)) return MPTV_ERR_ARG;
  const uint32_t k0 = in->key_off[i0], k1 = in->key_off[i1];
  if (!g.up) CK(cudaEventCreateWithFlags(&g.up, cudaEventDisableTiming));
  CK(g.h_koff.reserve(4 * (ni + 1)));
  CK(g.h_voff.reserve(8 * ni + 8));
  CK(g.h_tfirst.reserve(4 * (nt + 1)));
  CK(g.h_vlen.reserve(4 * ni + 4));
  CK(g.h_keys.reserve((size_t)(k1 - k0) + 16));
  uint32_t* koff = static_cast<uint32_t*>(g.h_koff.p);
  uint64_t* voff = static_cast<uint64_t*>(g.h_voff.p);
  uint32_t* tfirst = static_cast<uint32_t*>(g.h_tfirst.p);
  for (uint64_t i = 0; i <= ni; i++) koff[i] = in->key_off[i0 + i] - k0;
  for (uint64_t i = 0; i < ni; i++) {
    const uint64_t o = in->value_off[i0 + i];
    if (o & 15) return MPTV_ERR_ALIGN;
    if (o < b0 || o + in->value_len[i0 + i] > b1) return MPTV_ERR_ARG;
    voff[i] = o - b0;
  }
  for (uint64_t t = 0; t <= nt; t++) tfirst[t] = in->trie_first[cs + t] - i0;
  CK(g.key_bytes.reserve((size_t)(k1 - k0) + 16));
  CK(g.key_off.reserve(4 * (ni + 1)));
  CK(g.value_bytes.reserve((size_t)(b1 - b0) + 16));
  CK(g.value_off.reserve(8 * ni + 8));
  CK(g.value_len.reserve(4 * ni + 4));
  CK(g.trie_first.reserve(4 * (nt + 1)));
  // every small array goes through page-locked staging (a copy from pageable memory would block this
  // thread until the stream has drained, i.e. behind the previous chunk's gigabyte of values) and is
  // queued BEFORE the value bytes, which are copied straight out of the caller's arena
  if (k1 > k0) {
    memcpy(g.h_keys.p, in->key_bytes + k0, k1 - k0);
    CK(cudaMemcpyAsync(g.key_bytes.p, g.h_keys.p, k1 - k0, cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemcpyAsync(g.key_off.p, koff, 4 * (ni + 1), cudaMemcpyHostToDevice, st));
  if (ni) {
    memcpy(g.h_vlen.p, in->value_len + i0, 4 * ni);
    CK(cudaMemcpyAsync(g.value_off.p, voff, 8 * ni, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g.value_len.p, g.h_vlen.p, 4 * ni, cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemcpyAsync(g.trie_first.p, tfirst, 4 * (nt + 1), cudaMemcpyHostToDevice, st));
  const uint64_t copy_end = b1 < in->value_bytes_len ? b1 : in->value_bytes_len;
  if (copy_end > b0) CK(cudaMemcpyAsync(g.value_bytes.p, in->value_bytes + b0, copy_end - b0, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(g.up, st));
  b.key_bytes = g.key_bytes.as<uint8_t>(); b.key_off = g.key_off.as<uint32_t>();
  b.value_bytes = g.value_bytes.as<uint8_t>(); b.value_off = g.value_off.as<uint64_t>();
  b.value_len = g.value_len.as<uint32_t>(); b.trie_first = g.trie_first.as<uint32_t>();
  b.n_tries = (uint32_t)nt; b.n_items = ni;
  return MPTV_OK;
}

// end of the chunk of tries that starts at cs: about `chunk` value bytes
uint64_t kv_chunk_end(const mptv_kv_batch* in, uint64_t cs, uint64_t t1, uint64_t chunk) {
  uint64_t ce = cs + 1;
  const uint32_t i0 = in->trie_first[cs];
  const uint64_t b0 = i0 < in->n_items ? in->value_off[i0] : in->value_bytes_len;
  while (ce < t1) {
    const uint32_t ie = in->trie_first[ce + 1];
    const uint64_t be = ie < in->n_items ? in->value_off[ie] : in->value_bytes_len;
    if (be - b0 > chunk) break;
    ce++;
  }
  return ce;
}

// one device's share [t0, t1) of a host batch: chunks of about 1 GiB of values, the copy of chunk c+1
// (copy stream, second input stage) overlapping the rebuild of chunk c
int rebuild_slice_run(mptv_ctx* ctx, Device& d, const mptv_kv_batch* in, uint8_t* roots32, uint64_t t0, uint64_t t1) {
  if (t1 <= t0) return MPTV_OK;
  CK(cudaSetDevice(d.id));
  Rebuild& rb = d.rb;
  cudaStream_t st = d.stream, cp = d.slot[0].stream;
  const uint64_t chunk = 1ull << 30;
  TrieBatchDev cur, nxt;
  uint64_t cs = t0, ce = kv_chunk_end(in, cs, t1, chunk);
  int rc = upload_kv(ctx, d, in, cs, ce, rb.stage[0], cur, cp);
  if (rc != MPTV_OK) return rc;
  for (int k = 0; cs < t1; k ^= 1) {
    const uint64_t ns = ce, ne = ns < t1 ? kv_chunk_end(in, ns, t1, chunk) : ns;
    if (ns < t1) {  // stage k^1 is free: the rebuild that read it was synchronised at the end of the last iteration
      rc = upload_kv(ctx, d, in, ns, ne, rb.stage[k ^ 1], nxt, cp);
      if (rc != MPTV_OK) return rc;
    }
    const uint64_t nt = ce - cs;
    CK(cudaStreamWaitEvent(st, rb.stage[k].up, 0));
    CK(rb.out_roots.reserve(32 * nt));
    CK(rb.h_roots.reserve(32 * nt));
    rc = rebuild_on_device(ctx, d, cur, rb.out_roots.as<uint8_t>(), st);
    if (rc != MPTV_OK) return rc;
    CK(cudaMemcpyAsync(rb.h_roots.p, rb.out_roots.p, 32 * nt, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(roots32 + 32 * cs, rb.h_roots.p, 32 * nt);
    cs = ns; ce = ne; cur = nxt;
  }
  return MPTV_OK;
}

int rebuild_slice(mptv_ctx* ctx, Device& d, const mptv_kv_batch* in, uint8_t* roots32, uint64_t t0, uint64_t t1) {
  const int rc = rebuild_slice_run(ctx, d, in, roots32, t0, t1);
  if (rc != MPTV_OK) quiesce(d);  // the next chunk's copy may still be reading the caller's arena
  return rc;
}

int check_kv(const mptv_kv_batch* in, const uint8_t* roots32) {
  if (!roots32 || !in->trie_first || !in->key_off || in->n_tries > 0x7fffffffull || in->n_items > 0x50000000ull)
    return MPTV_ERR_ARG;
  if (in->n_items && (!in->key_bytes || !in->value_bytes || !in->value_off || !in->value_len)) return MPTV_ERR_ARG;
  if (in->trie_first[in->n_tries] > in->n_items) return MPTV_ERR_ARG;
  for (uint64_t t = 0; t < in->n_tries; t++)
    if (in->trie_first[t + 1] < in->trie_first[t]) return MPTV_ERR_ARG;
  for (uint64_t i = 0; i < in->n_items; i++)
    if (in->key_off[i + 1] < in->key_off[i]) return MPTV_ERR_ARG;
  for (uint64_t i = 0; i + 1 < in->n_items; i++)
    if (in->value_off[i + 1] < in->value_off[i] + in->value_len[i]) return MPTV_ERR_ARG;  // laid out in item order
  return MPTV_OK;
}

}  // namespace

extern "C" {

int mptv_trie_roots_device(mptv_ctx* ctx, int dev_index, const mptv_kv_batch* in, uint8_t* roots32, void* stream) {
  if (!ctx || !in || dev_index < 0 || dev_index >= (int)ctx->dev.size()) return MPTV_ERR_ARG;
  if (in->n_tries == 0) return MPTV_OK;
  if (!roots32 || !in->trie_first || !in->key_off || in->n_tries > 0x7fffffffull || in->n_items > 0x50000000ull)
    return MPTV_ERR_ARG;
  if (in->n_items && (!in->key_bytes || !in->value_bytes || !in->value_off || !in->value_len)) return MPTV_ERR_ARG;
  Device& d = ctx->dev[dev_index];
  CK(cudaSetDevice(d.id));
  cudaStream_t st = stream ? (cudaStream_t)stream : d.stream;
  TrieBatchDev b;
  b.key_bytes = in->key_bytes; b.key_off = in->key_off; b.value_bytes = in->value_bytes; b.value_off = in->value_off;
  b.value_len = in->value_len; b.trie_first = in->trie_first; b.n_tries = (uint32_t)in->n_tries; b.n_items = in->n_items;
  return rebuild_on_device(ctx, d, b, roots32, st);
}

int mptv_trie_roots(mptv_ctx* ctx, const mptv_kv_batch* in, uint8_t* roots32) {
  if (!ctx || !in) return MPTV_ERR_ARG;
  if (in->n_tries == 0) return MPTV_OK;
  const int ck = check_kv(in, roots32);
  if (ck != MPTV_OK) return ck;
  const int nd = (int)ctx->dev.size();
  std::vector<uint64_t> cut(nd + 1, 0);
  cut[nd] = in->n_tries;
  if (nd > 1) {
    uint64_t t = 0;
    for (int k = 1; k < nd; k++) {
      const uint64_t target = in->value_bytes_len / nd * k;
      uint64_t lo = t, hi = in->n_tries;
      while (lo < hi) {
        const uint64_t mid = (lo + hi) / 2;
        const uint32_t fi = in->trie_first[mid];
        const uint64_t off = fi < in->n_items ? in->value_off[fi] : in->value_bytes_len;
        if (off < target) lo = mid + 1; else hi = mid;
      }
      cut[k] = t = lo;
    }
  }
  std::vector<int> rcs(nd, MPTV_OK);
  if (nd == 1) rcs[0] = rebuild_slice(ctx, ctx->dev[0], in, roots32, cut[0], cut[1]);
  else {
    std::vector<std::thread> th;
    for (int k = 0; k < nd; k++)
      th.emplace_back([&, k] { rcs[k] = rebuild_slice(ctx, ctx->dev[k], in, roots32, cut[k], cut[k + 1]); });
    for (auto& x : th) x.join();
  }
  for (int k = 0; k < nd; k++) if (rcs[k] != MPTV_OK) return rcs[k];
  return MPTV_OK;
}

// One device's share of mptv_trie_proofs: tries [t0, t1) and the targets [q0, q1) that look into them.
struct ProofSlice {
  uint64_t t0 = 0, t1 = 0, q0 = 0, q1 = 0;
  uint32_t n_nodes = 0;   // phase 1 out: nodes / padded bytes of the slice's proofs
  uint64_t n_bytes = 0;
  TrieBatchDev b;
  int rc = MPTV_OK;
};

// phase 1: rebuild the slice's tries, hand back their roots, count the proof nodes of its targets
static int trie_proofs_count(mptv_ctx* ctx, Device& d, const mptv_kv_batch* in, const mptv_proof_targets* tg, uint8_t* roots32,
                             ProofSlice& sl) {
  Rebuild& rb = d.rb;
  CK(cudaSetDevice(d.id));
  cudaStream_t st = d.stream;
  const uint64_t nt = sl.t1 - sl.t0, nq = sl.q1 - sl.q0;
  if (nt == 0) return MPTV_OK;
  int rc = upload_kv(ctx, d, in, sl.t0, sl.t1, rb.stage[0], sl.b, st);
  if (rc != MPTV_OK) return rc;
  CK(rb.out_roots.reserve(32 * nt));
  rc = rebuild_on_device(ctx, d, sl.b, rb.out_roots.as<uint8_t>(), st);
  if (rc != MPTV_OK) return rc;
  CK(cudaMemcpyAsync(roots32 + 32 * sl.t0, rb.out_roots.p, 32 * nt, cudaMemcpyDeviceToHost, st));
  if (nq == 0) { CK(cudaStreamSynchronize(st)); return MPTV_OK; }
  TrieWork w;
  w.rec = rb.rec.as<uint4>(); w.off = rb.off.as<uint64_t>(); w.len = rb.len.as<uint32_t>();
  w.digests = rb.digests.as<uint8_t>(); w.tcount = rb.tcount.as<uint32_t>(); w.lvl_list = rb.lvl_list.as<uint32_t>();
  w.sum = rb.sum.as<TrieSummary>();
  const uint32_t k0 = tg->key_off[sl.q0], kb = tg->key_off[sl.q1] - k0;
  // trie indices and key offsets relative to the slice
  std::vector<uint32_t> ltrie(nq), lkoff(nq + 1);
  for (uint64_t q = 0; q < nq; q++) ltrie[q] = tg->trie[sl.q0 + q] - (uint32_t)sl.t0;
  for (uint64_t q = 0; q <= nq; q++) lkoff[q] = tg->key_off[sl.q0 + q] - k0;
  CK(rb.q_trie.reserve(4 * nq));
  CK(rb.q_key_bytes.reserve((size_t)kb + 16));
  CK(rb.q_key_off.reserve(4 * (nq + 1)));
  CK(rb.q_cnt.reserve(4 * nq));
  CK(rb.q_bytes.reserve(8 * nq));
  CK(rb.q_proof_first.reserve(4 * (nq + 1)));
  CK(rb.q_byte_first.reserve(8 * (nq + 1)));
  CK(cudaMemcpyAsync(rb.q_trie.p, ltrie.data(), 4 * nq, cudaMemcpyHostToDevice, st));
  if (kb) CK(cudaMemcpyAsync(rb.q_key_bytes.p, tg->key_bytes + k0, kb, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(rb.q_key_off.p, lkoff.data(), 4 * (nq + 1), cudaMemcpyHostToDevice, st));
  CK(launch_trie_proof_count(sl.b, w, rb.q_trie.as<uint32_t>(), rb.q_key_bytes.as<uint8_t>(), rb.q_key_off.as<uint32_t>(),
                             (uint32_t)nq, rb.q_cnt.as<uint32_t>(), rb.q_bytes.as<uint64_t>(),
                             rb.q_proof_first.as<uint32_t>(), rb.q_byte_first.as<uint64_t>(), st));
  CK(cudaMemcpyAsync(&sl.n_nodes, rb.q_proof_first.as<uint32_t>() + nq, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&sl.n_bytes, rb.q_byte_first.as<uint64_t>() + nq, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));  // (the pageable copies above were staged by the driver before returning)
  return MPTV_OK;
}

// phase 2: emit the slice's proofs and copy them to their place in the caller's arrays
static int trie_proofs_emit(mptv_ctx* ctx, Device& d, const ProofSlice& sl, mptv_proofs_out* out, uint64_t node_base, uint64_t byte_base) {
  const uint64_t nq = sl.q1 - sl.q0;
  if (nq == 0) return MPTV_OK;
  Rebuild& rb = d.rb;
  CK(cudaSetDevice(d.id));
  cudaStream_t st = d.stream;
  TrieWork w;
  w.rec = rb.rec.as<uint4>(); w.off = rb.off.as<uint64_t>(); w.len = rb.len.as<uint32_t>();
  w.digests = rb.digests.as<uint8_t>(); w.tcount = rb.tcount.as<uint32_t>(); w.lvl_list = rb.lvl_list.as<uint32_t>();
  w.sum = rb.sum.as<TrieSummary>();
  CK(rb.q_out_bytes.reserve((size_t)sl.n_bytes + 16));
  CK(rb.q_out_off.reserve(8 * (size_t)sl.n_nodes + 8));
  CK(rb.q_out_len.reserve(4 * (size_t)sl.n_nodes + 4));
  CK(cudaMemsetAsync(rb.q_out_bytes.p, 0, (size_t)sl.n_bytes + 16, st));
  CK(launch_trie_proof_emit(sl.b, w, rb.arena.as<uint8_t>(), rb.q_trie.as<uint32_t>(), rb.q_key_bytes.as<uint8_t>(),
                            rb.q_key_off.as<uint32_t>(), (uint32_t)nq, rb.q_proof_first.as<uint32_t>(),
                            rb.q_byte_first.as<uint64_t>(), rb.q_out_bytes.as<uint8_t>(), rb.q_out_off.as<uint64_t>(),
                            rb.q_out_len.as<uint32_t>(), rb.leaves_in_arena, st));
  if (sl.n_bytes) CK(cudaMemcpyAsync(out->node_bytes + byte_base, rb.q_out_bytes.p, (size_t)sl.n_bytes, cudaMemcpyDeviceToHost, st));
  if (sl.n_nodes) {
    CK(cudaMemcpyAsync(out->node_off + node_base, rb.q_out_off.p, 8 * (size_t)sl.n_nodes, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out->node_len + node_base, rb.q_out_len.p, 4 * (size_t)sl.n_nodes, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaMemcpyAsync(out->proof_first + sl.q0, rb.q_proof_first.p, 4 * nq, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  // the slice's offsets and node indices were relative to the slice: move them to their place in the whole output
  if (byte_base) for (uint64_t i = 0; i < sl.n_nodes; i++) out->node_off[node_base + i] += byte_base;
  if (node_base) for (uint64_t q = 0; q < nq; q++) out->proof_first[sl.q0 + q] += (uint32_t)node_base;
  return MPTV_OK;
}

static int trie_proofs_run(mptv_ctx* ctx, const mptv_kv_batch* in, const mptv_proof_targets* tg, uint8_t* roots32,
                           mptv_proofs_out* out) {
  if (!ctx || !in || !tg || !out) return MPTV_ERR_ARG;
  out->n_nodes = 0; out->node_bytes_len = 0;
  const uint64_t nq = tg->n_targets;
  if (nq > 0xfffffff0ull || (nq && (!tg->trie || !tg->key_off || !out->proof_first))) return MPTV_ERR_ARG;
  if (in->n_tries == 0) return nq ? MPTV_ERR_ARG : MPTV_OK;
  const int ck = check_kv(in, roots32);
  if (ck != MPTV_OK) return ck;
  bool ordered = true;  // targets grouped by trie in ascending order: the usual shape, and what lets the devices share them
  for (uint64_t q = 0; q < nq; q++) {
    if (tg->trie[q] >= in->n_tries || tg->key_off[q + 1] < tg->key_off[q] ||
        tg->key_off[q + 1] - tg->key_off[q] > (uint32_t)kTrieMaxKeyLen)
      return MPTV_ERR_ARG;
    if (q && tg->trie[q] < tg->trie[q - 1]) ordered = false;
  }
  // tries are cut into contiguous slices with equal shares of the value bytes, one per device (no inter-device
  // traffic); the targets of a slice are then a contiguous range of the (trie-ordered) target list
  int nd = (int)ctx->dev.size();
  if (!ordered || in->n_tries < 8ull * nd) nd = 1;
  std::vector<ProofSlice> sl(nd);
  {
    uint64_t t = 0, q = 0;
    for (int k = 0; k < nd; k++) {
      sl[k].t0 = t;
      if (k == nd - 1) t = in->n_tries;
      else {
        const uint64_t target = in->value_bytes_len / nd * (k + 1);
        uint64_t lo = t, hi = in->n_tries;
        while (lo < hi) {
          const uint64_t mid = (lo + hi) / 2;
          const uint32_t fi = in->trie_first[mid];
          const uint64_t off = fi < in->n_items ? in->value_off[fi] : in->value_bytes_len;
          if (off < target) lo = mid + 1; else hi = mid;
        }
        t = lo;
      }
      sl[k].t1 = t;
      sl[k].q0 = q;
      if (nd == 1) q = nq;
      else while (q < nq && tg->trie[q] < t) q++;
      sl[k].q1 = q;
    }
  }
  auto for_each_device = [&](auto&& fn) {
    if (nd == 1) { sl[0].rc = fn(0); return; }
    std::vector<std::thread> th;
    for (int k = 0; k < nd; k++) th.emplace_back([&, k] { sl[k].rc = fn(k); });
    for (auto& x : th) x.join();
  };
  for_each_device([&](int k) { return trie_proofs_count(ctx, ctx->dev[k], in, tg, roots32, sl[k]); });
  for (int k = 0; k < nd; k++) if (sl[k].rc != MPTV_OK) return sl[k].rc;
  if (nq == 0) return MPTV_OK;
  std::vector<uint64_t> node_base(nd + 1, 0), byte_base(nd + 1, 0);
  for (int k = 0; k < nd; k++) { node_base[k + 1] = node_base[k] + sl[k].n_nodes; byte_base[k + 1] = byte_base[k] + sl[k].n_bytes; }
  out->n_nodes = node_base[nd];
  out->node_bytes_len = byte_base[nd] + 16;
  if (node_base[nd] > 0xfffffff0ull) return MPTV_ERR_ARG;
  if (node_base[nd] > out->nodes_cap || byte_base[nd] + 16 > out->node_bytes_cap || !out->node_bytes || !out->node_off ||
      !out->node_len)
    return MPTV_ERR_NOMEM;  // n_nodes / node_bytes_len hold what is required
  for_each_device([&](int k) { return trie_proofs_emit(ctx, ctx->dev[k], sl[k], out, node_base[k], byte_base[k]); });
  for (int k = 0; k < nd; k++) if (sl[k].rc != MPTV_OK) return sl[k].rc;
  out->proof_first[nq] = (uint32_t)node_base[nd];
  memset(out->node_bytes + byte_base[nd], 0, 16);
  return MPTV_OK;
}

int mptv_trie_proofs(mptv_ctx* ctx, const mptv_kv_batch* in, const mptv_proof_targets* tg, uint8_t* roots32,
                     mptv_proofs_out* out) {
  int rc;
  try {
    rc = trie_proofs_run(ctx, in, tg, roots32, out);
  } catch (...) {
    rc = MPTV_ERR_NOMEM;
  }
  if (rc != MPTV_OK && rc != MPTV_ERR_NOMEM && ctx) for (Device& d : ctx->dev) quiesce(d);
  return rc;
}

int mptv_last_rebuild_timings(mptv_ctx* ctx, int dev_index, mptv_rebuild_timings* out) {
  if (!ctx || !out || dev_index < 0 || dev_index >= (int)ctx->dev.size()) return MPTV_ERR_ARG;
  Device& d = ctx->dev[dev_index];
  Rebuild& rb = d.rb;
  memset(out, 0, sizeof *out);
  if (!rb.have_timing) return MPTV_ERR_ARG;
  CK(cudaSetDevice(d.id));
  CK(cudaEventSynchronize(rb.ev_end));
  CK(cudaEventElapsedTime(&out->structure_ms, rb.ev_begin, rb.ev_struct));
  CK(cudaEventElapsedTime(&out->total_ms, rb.ev_begin, rb.ev_end));
  for (uint32_t h = 0; h < rb.levels; h++) {
    float a = 0, b = 0;
    CK(cudaEventElapsedTime(&a, rb.lvl_ev[3 * h], rb.lvl_ev[3 * h + 1]));
    CK(cudaEventElapsedTime(&b, rb.lvl_ev[3 * h + 1], rb.lvl_ev[3 * h + 2]));
    out->encode_ms += a;
    out->keccak_ms += b;
  }
  out->n_nodes = rb.n_nodes; out->n_hashed = rb.n_hashed; out->n_perm = rb.n_perm; out->arena_bytes = rb.arena_bytes;
  out->levels = rb.levels; out->keccak_launches = rb.keccak_launches; out->other_launches = rb.other_launches;
  return MPTV_OK;
}

}  // extern "C"
