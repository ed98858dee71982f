// workload/mptgen.cpp -- synthetic workload generator (host C++, test/bench infrastructure).
//
// Stands in for the network half of trie-utils (eth_getProof over RPC,
// /root/reference/trie-utils/src/proofs/account.rs:24-74, storage.rs:24-121): builds synthetic
// Ethereum state / storage tries of the shapes BASELINE.json names (SURVEY.md section 8d configs
// 2, 3, 5) and cuts account / storage proofs out of them straight into the CSR arena of
// include/mptv.h, with the config-3 mutators (bit flips, dropped nodes, wrong key, shuffle, junk).
// It is NOT part of the verification path: the product never calls it, and it shares no code with
// oracle/.  Its own CPU Keccak exists only to hash the trie it fabricates.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------ keccak-256 (generator only)
const uint64_t RC[24] = {
  0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
  0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
  0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
  0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
  0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
  0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
#define ROL(x,n) (((x)<<(n))|((x)>>(64-(n))))
void keccakf(uint64_t a[25]) {
  uint64_t a0=a[0],a1=a[1],a2=a[2],a3=a[3],a4=a[4],a5=a[5],a6=a[6],a7=a[7],a8=a[8],a9=a[9],a10=a[10],a11=a[11],a12=a[12],a13=a[13],a14=a[14],a15=a[15],a16=a[16],a17=a[17],a18=a[18],a19=a[19],a20=a[20],a21=a[21],a22=a[22],a23=a[23],a24=a[24];
  for (int r = 0; r < 24; r++) {
    uint64_t c0=a0^a5^a10^a15^a20, c1=a1^a6^a11^a16^a21, c2=a2^a7^a12^a17^a22, c3=a3^a8^a13^a18^a23, c4=a4^a9^a14^a19^a24;
    uint64_t d0=c4^ROL(c1,1), d1=c0^ROL(c2,1), d2=c1^ROL(c3,1), d3=c2^ROL(c4,1), d4=c3^ROL(c0,1);
    uint64_t b0=a0^d0, b10=ROL(a1^d1,1), b20=ROL(a2^d2,62), b5=ROL(a3^d3,28), b15=ROL(a4^d4,27);
    uint64_t b16=ROL(a5^d0,36), b1=ROL(a6^d1,44), b11=ROL(a7^d2,6), b21=ROL(a8^d3,55), b6=ROL(a9^d4,20);
    uint64_t b7=ROL(a10^d0,3), b17=ROL(a11^d1,10), b2=ROL(a12^d2,43), b12=ROL(a13^d3,25), b22=ROL(a14^d4,39);
    uint64_t b23=ROL(a15^d0,41), b8=ROL(a16^d1,45), b18=ROL(a17^d2,15), b3=ROL(a18^d3,21), b13=ROL(a19^d4,8);
    uint64_t b14=ROL(a20^d0,18), b24=ROL(a21^d1,2), b9=ROL(a22^d2,61), b19=ROL(a23^d3,56), b4=ROL(a24^d4,14);
    a0=b0^(~b1&b2)^RC[r]; a1=b1^(~b2&b3); a2=b2^(~b3&b4); a3=b3^(~b4&b0); a4=b4^(~b0&b1);
    a5=b5^(~b6&b7); a6=b6^(~b7&b8); a7=b7^(~b8&b9); a8=b8^(~b9&b5); a9=b9^(~b5&b6);
    a10=b10^(~b11&b12); a11=b11^(~b12&b13); a12=b12^(~b13&b14); a13=b13^(~b14&b10); a14=b14^(~b10&b11);
    a15=b15^(~b16&b17); a16=b16^(~b17&b18); a17=b17^(~b18&b19); a18=b18^(~b19&b15); a19=b19^(~b15&b16);
    a20=b20^(~b21&b22); a21=b21^(~b22&b23); a22=b22^(~b23&b24); a23=b23^(~b24&b20); a24=b24^(~b20&b21);
  }
  a[0]=a0;a[1]=a1;a[2]=a2;a[3]=a3;a[4]=a4;a[5]=a5;a[6]=a6;a[7]=a7;a[8]=a8;a[9]=a9;a[10]=a10;a[11]=a11;a[12]=a12;a[13]=a13;a[14]=a14;a[15]=a15;a[16]=a16;a[17]=a17;a[18]=a18;a[19]=a19;a[20]=a20;a[21]=a21;a[22]=a22;a[23]=a23;a[24]=a24;
}
void keccak256(const uint8_t* in, size_t len, uint8_t out[32]) {
  uint64_t s[25] = {0};
  while (len >= 136) {
    for (int i = 0; i < 17; i++) { uint64_t w; memcpy(&w, in + 8 * i, 8); s[i] ^= w; }
    keccakf(s); in += 136; len -= 136;
  }
  uint8_t last[136] = {0};
  memcpy(last, in, len);
  last[len] ^= 0x01; last[135] ^= 0x80;
  for (int i = 0; i < 17; i++) { uint64_t w; memcpy(&w, last + 8 * i, 8); s[i] ^= w; }
  keccakf(s);
  memcpy(out, s, 32);
}

// ------------------------------------------------------------------ rng
struct SplitMix {
  uint64_t s;
  explicit SplitMix(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  void fill(uint8_t* p, size_t n) {
    while (n >= 8) { uint64_t v = next(); memcpy(p, &v, 8); p += 8; n -= 8; }
    if (n) { uint64_t v = next(); memcpy(p, &v, n); }
  }
};
inline uint64_t mix(uint64_t a, uint64_t b) { SplitMix m(a * 0x9e3779b97f4a7c15ULL + b + 0x632be59bd9b4e019ULL); m.next(); return m.next(); }

// ------------------------------------------------------------------ rlp helpers
struct Buf {
  uint8_t d[640];
  uint32_t n = 0;
  void put(uint8_t b) { d[n++] = b; }
  void put(const uint8_t* p, uint32_t k) { memcpy(d + n, p, k); n += k; }
};
inline uint32_t rlp_hdr(uint8_t* out, uint32_t n, bool list) {
  uint8_t base = list ? 0xC0 : 0x80;
  if (n < 56) { out[0] = (uint8_t)(base + n); return 1; }
  if (n < 256) { out[0] = (uint8_t)(base + 56); out[1] = (uint8_t)n; return 2; }
  out[0] = (uint8_t)(base + 57); out[1] = (uint8_t)(n >> 8); out[2] = (uint8_t)n; return 3;
}
inline void rlp_str(Buf& b, const uint8_t* p, uint32_t n) {
  if (n == 1 && p[0] < 0x80) { b.put(p[0]); return; }
  b.n += rlp_hdr(b.d + b.n, n, false);
  b.put(p, n);
}
inline void rlp_uint_be(Buf& b, const uint8_t* be, uint32_t n) {  // big-endian integer, strip zeros
  while (n && *be == 0) { be++; n--; }
  if (n == 0) { b.put(0x80); return; }
  rlp_str(b, be, n);
}

typedef std::array<uint8_t, 32> H256;

const uint8_t EMPTY_ROOT[32] = {0x56, 0xe8, 0x1f, 0x17, 0x1b, 0xcc, 0x55, 0xa6, 0xff, 0x83, 0x45, 0xe6, 0x92, 0xc0, 0xf8, 0x6e,
                                0x5b, 0x48, 0xe0, 0x1b, 0x99, 0x6c, 0xad, 0xc0, 0x01, 0x62, 0x2f, 0xb5, 0xe3, 0x63, 0xb4, 0x21};

// ------------------------------------------------------------------ trie
enum { K_LEAF = 0, K_EXT = 1, K_BRANCH = 2 };
const uint32_t NONE = 0xffffffffu;

struct Node {
  uint64_t off;
  uint32_t len;
  uint32_t child0;  // index into children[] (16 entries for a branch, 1 for an extension)
  uint8_t kind;
  uint8_t plen;     // path nibbles of a leaf / extension
};

struct Part {  // output of one build task
  std::vector<uint8_t> arena;
  std::vector<Node> nodes;
  std::vector<uint32_t> children;
  std::vector<H256> hashes;  // parallel to nodes (valid when len >= 32)
};

struct Trie {
  int kind = 0;  // 0 account values, 1 storage values
  uint64_t seed = 0;
  std::vector<H256> keys;        // sorted
  std::vector<uint32_t> key_src; // sorted position -> original index
  std::vector<uint32_t> src_pos; // original index -> sorted position
  std::vector<H256> pool_roots;  // storage roots handed to accounts (kind 0), may be empty
  Part p;
  uint32_t root = NONE;
  H256 root_hash;
  // storage tries (kind 1): every proof key has a pre-image, as the wire form of the nested workload needs
  // (StorageProofInput carries RAW storage keys, the guest hashes them): the key of an absent / wrong-key proof is
  // keccak(32 random bytes) instead of 32 random bytes -- the same distribution
  bool raw_keys = false;
};

inline int nib(const H256& k, int i) { return (i & 1) ? (k[i >> 1] & 15) : (k[i >> 1] >> 4); }

// value of the key with ORIGINAL index i
void make_value(const Trie& t, uint32_t i, Buf& out) {
  SplitMix r(mix(t.seed ^ 0x5a5a, i));
  if (t.kind == 0) {
    // rlp([nonce u64 < 2^20, balance u128, storage_root, code_hash])   
    Buf pl;
    uint64_t nonce = r.next() & 0xfffff;
    uint8_t be[16];
    for (int k = 0; k < 8; k++) be[k] = (uint8_t)(nonce >> (56 - 8 * k));
    rlp_uint_be(pl, be, 8);
    uint64_t b0 = r.next(), b1 = r.next();
    if ((b0 & 7) == 0) b0 = 0;               // some balances fit 64 bits
    for (int k = 0; k < 8; k++) { be[k] = (uint8_t)(b0 >> (56 - 8 * k)); be[8 + k] = (uint8_t)(b1 >> (56 - 8 * k)); }
    rlp_uint_be(pl, be, 16);
    uint8_t h[32];
    if (!t.pool_roots.empty()) memcpy(h, t.pool_roots[i % t.pool_roots.size()].data(), 32);
    else if (r.next() & 1) memcpy(h, EMPTY_ROOT, 32);
    else r.fill(h, 32);
    rlp_str(pl, h, 32);
    r.fill(h, 32);
    rlp_str(pl, h, 32);
    out.n = rlp_hdr(out.d, pl.n, true);
    out.put(pl.d, pl.n);
  } else {
    // rlp(u256 balance); 5 % of balances in 1..0x7f (single-byte path of rule R20)
    uint8_t be[32];
    uint64_t sel = r.next() % 100;
    if (sel < 5) { out.put((uint8_t)(1 + r.next() % 0x7f)); return; }
    uint32_t n = 1 + (uint32_t)(r.next() % 12);
    r.fill(be, n);
    if (be[0] == 0) be[0] = 1;
    if (n == 1 && be[0] < 0x80) be[0] |= 0x80;
    rlp_str(out, be, n);
  }
}

inline void hex_prefix(Buf& b, const H256& key, int from, int cnt, bool leaf) {
  uint8_t tmp[33];
  uint32_t k = 0;
  uint8_t flag = leaf ? 0x20 : 0x00;
  int i = from;
  if (cnt & 1) { tmp[k++] = (uint8_t)(flag | 0x10 | nib(key, i)); i++; }
  else tmp[k++] = flag;
  for (; i < from + cnt; i += 2) tmp[k++] = (uint8_t)((nib(key, i) << 4) | nib(key, i + 1));
  rlp_str(b, tmp, k);
}

struct Builder {
  const Trie& t;
  Part& p;
  Builder(const Trie& tt, Part& pp) : t(tt), p(pp) {}

  uint32_t add(const Buf& enc, uint8_t kind, uint8_t plen, uint32_t child0) {
    Node n;
    n.off = p.arena.size(); n.len = enc.n; n.kind = kind; n.plen = plen; n.child0 = child0;
    p.arena.insert(p.arena.end(), enc.d, enc.d + enc.n);
    H256 h;
    if (enc.n >= 32) keccak256(enc.d, enc.n, h.data()); else h.fill(0);
    p.nodes.push_back(n);
    p.hashes.push_back(h);
    return (uint32_t)p.nodes.size() - 1;
  }
  void put_ref(Buf& b, uint32_t id) {
    const Node& n = p.nodes[id];
    if (n.len < 32) b.put(p.arena.data() + n.off, n.len);
    else rlp_str(b, p.hashes[id].data(), 32);
  }
  // keys [lo, hi) share their first `depth` nibbles
  uint32_t build(uint32_t lo, uint32_t hi, int depth) {
    if (hi - lo == 1) {
      Buf pl, val, enc;
      hex_prefix(pl, t.keys[lo], depth, 64 - depth, true);
      make_value(t, t.key_src[lo], val);
      rlp_str(pl, val.d, val.n);
      enc.n = rlp_hdr(enc.d, pl.n, true);
      enc.put(pl.d, pl.n);
      return add(enc, K_LEAF, (uint8_t)(64 - depth), NONE);
    }
    int cp = 0;
    while (depth + cp < 64 && nib(t.keys[lo], depth + cp) == nib(t.keys[hi - 1], depth + cp)) cp++;
    if (cp > 0) {
      uint32_t c = build(lo, hi, depth + cp);
      Buf pl, enc;
      hex_prefix(pl, t.keys[lo], depth, cp, false);
      put_ref(pl, c);
      enc.n = rlp_hdr(enc.d, pl.n, true);
      enc.put(pl.d, pl.n);
      uint32_t c0 = (uint32_t)p.children.size();
      p.children.push_back(c);
      return add(enc, K_EXT, (uint8_t)cp, c0);
    }
    uint32_t kids[16];
    uint32_t s = lo;
    for (int v = 0; v < 16; v++) {
      uint32_t e = s;
      while (e < hi && nib(t.keys[e], depth) == v) e++;
      kids[v] = e > s ? build(s, e, depth + 1) : NONE;
      s = e;
    }
    return finish_branch(kids);
  }
  uint32_t finish_branch(const uint32_t kids[16]) {
    Buf pl, enc;
    for (int v = 0; v < 16; v++) {
      if (kids[v] == NONE) pl.put(0x80); else put_ref(pl, kids[v]);
    }
    pl.put(0x80);
    enc.n = rlp_hdr(enc.d, pl.n, true);
    enc.put(pl.d, pl.n);
    uint32_t c0 = (uint32_t)p.children.size();
    for (int v = 0; v < 16; v++) p.children.push_back(kids[v]);
    return add(enc, K_BRANCH, 0, c0);
  }
};

void append_part(Part& dst, Part& src, uint32_t& node_base) {
  node_base = (uint32_t)dst.nodes.size();
  const uint64_t ab = dst.arena.size();
  const uint32_t cb = (uint32_t)dst.children.size();
  dst.arena.insert(dst.arena.end(), src.arena.begin(), src.arena.end());
  for (Node n : src.nodes) {
    n.off += ab;
    if (n.child0 != NONE) n.child0 += cb;
    dst.nodes.push_back(n);
  }
  for (uint32_t c : src.children) dst.children.push_back(c == NONE ? NONE : c + node_base);
  dst.hashes.insert(dst.hashes.end(), src.hashes.begin(), src.hashes.end());
  Part().arena.swap(src.arena);
  std::vector<Node>().swap(src.nodes);
  std::vector<uint32_t>().swap(src.children);
  std::vector<H256>().swap(src.hashes);
}

void build_trie(Trie& t, int nthreads) {
  const uint32_t n = (uint32_t)t.keys.size();
  if (n == 0) { t.root = NONE; memcpy(t.root_hash.data(), EMPTY_ROOT, 32); return; }
  // bucket boundaries by first byte
  std::vector<uint32_t> b0(257, 0);
  for (uint32_t i = 0; i < n; i++) b0[t.keys[i][0] + 1]++;
  for (int i = 0; i < 256; i++) b0[i + 1] += b0[i];
  bool parallel = n >= 4096 && nthreads > 1;
  if (parallel) {
    // the two top levels must be plain branches: every first nibble needs >= 2 distinct second nibbles
    for (int a = 0; a < 16 && parallel; a++) {
      int distinct = 0;
      for (int c = 0; c < 16; c++) if (b0[16 * a + c + 1] > b0[16 * a + c]) distinct++;
      if (distinct < 2) parallel = false;
    }
  }
  if (!parallel) {
    Builder b(t, t.p);
    t.root = b.build(0, n, 0);
  } else {
    std::vector<Part> parts(256);
    std::vector<uint32_t> local_root(256, NONE);
    std::atomic<int> next(0);
    auto work = [&]() {
      for (;;) {
        int k = next.fetch_add(1);
        if (k >= 256) break;
        if (b0[k + 1] == b0[k]) continue;
        Builder b(t, parts[k]);
        local_root[k] = b.build(b0[k], b0[k + 1], 2);
      }
    };
    std::vector<std::thread> th;
    for (int i = 0; i < nthreads; i++) th.emplace_back(work);
    for (auto& x : th) x.join();
    std::vector<uint32_t> groot(256, NONE);
    size_t tot_nodes = 0, tot_arena = 0, tot_children = 0;
    for (auto& q : parts) { tot_nodes += q.nodes.size(); tot_arena += q.arena.size(); tot_children += q.children.size(); }
    t.p.nodes.reserve(tot_nodes + 32); t.p.arena.reserve(tot_arena + 16384);
    t.p.children.reserve(tot_children + 512); t.p.hashes.reserve(tot_nodes + 32);
    for (int k = 0; k < 256; k++) {
      if (local_root[k] == NONE) continue;
      uint32_t base;
      append_part(t.p, parts[k], base);
      groot[k] = local_root[k] + base;
    }
    Builder b(t, t.p);
    uint32_t top[16];
    for (int a = 0; a < 16; a++) top[a] = b.finish_branch(&groot[16 * a]);
    t.root = b.finish_branch(top);
  }
  const Node& r = t.p.nodes[t.root];
  keccak256(t.p.arena.data() + r.off, r.len, t.root_hash.data());
}

// key of ORIGINAL index i
// the storage slot of ORIGINAL index i: keccak(pad32(holder) || pad32(index)), the Solidity mapping layout
void make_slot(const Trie& t, uint64_t i, uint8_t slot[32]) {
  SplitMix r(mix(t.seed, i));
  uint8_t pre[64] = {0};
  r.fill(pre + 12, 20);
  pre[63] = (uint8_t)(t.seed & 7);
  keccak256(pre, 64, slot);
}

void make_key(const Trie& t, uint64_t i, H256& out) {
  if (t.kind == 0) {
    SplitMix r(mix(t.seed, i));
    uint8_t addr[20];
    r.fill(addr, 20);
    keccak256(addr, 20, out.data());  // account.rs:54  key = keccak(address)
  } else {
    uint8_t slot[32];
    make_slot(t, i, slot);
    keccak256(slot, 32, out.data());  // tests/storage.rs:78 digest_keccak(slot): trie key = keccak(slot)
  }
}

// ------------------------------------------------------------------ proofs
enum { MUT_NONE = 0, MUT_FLIP_LEAF = 1, MUT_FLIP_INNER = 2, MUT_DROP_LAST = 3, MUT_DROP_ROOT = 4,
       MUT_WRONG_KEY = 5, MUT_SHUFFLE = 6, MUT_JUNK = 7 };

struct Path { uint32_t ids[80]; uint32_t n = 0; };

// nodes on the path of `key`, root first; inline nodes are part of their parent
void walk(const Trie& t, const H256& key, Path& out) {
  out.n = 0;
  if (t.root == NONE) return;
  uint32_t cur = t.root;
  int depth = 0;
  for (;;) {
    const Node& n = t.p.nodes[cur];
    if (n.len >= 32 || cur == t.root) out.ids[out.n++] = cur;
    if (n.kind == K_LEAF) return;
    if (n.kind == K_EXT) {
      // compare the extension path (re-read from its encoding) with the key
      const uint8_t* e = t.p.arena.data() + n.off;
      uint32_t h = e[0] < 0xf8 ? 1 : 1 + (e[0] - 0xf7);
      const uint8_t* it = e + h;
      const uint8_t* hp = it[0] < 0x80 ? it : it + 1;
      int odd = (hp[0] >> 4) & 1;
      for (int k = 0; k < n.plen; k++) {
        int q = k + 2 - odd;
        int pn = (q & 1) ? (hp[q >> 1] & 15) : (hp[q >> 1] >> 4);
        if (pn != nib(key, depth + k)) return;
      }
      depth += n.plen;
      cur = t.p.children[n.child0];
      continue;
    }
    uint32_t c = t.p.children[n.child0 + nib(key, depth)];
    if (c == NONE) return;
    depth++;
    cur = c;
  }
}

// raw (may be null): the key's pre-image when the trie keeps one (Trie::raw_keys), else a copy of the key
void proof_key(const Trie& t, int64_t sel, uint64_t seed2, uint64_t i, H256& key, uint8_t* raw = nullptr) {
  const bool pre = t.raw_keys && t.kind == 1;
  if (sel >= 0) {
    key = t.keys[t.src_pos[(uint32_t)sel]];
    if (raw) { if (pre) make_slot(t, (uint64_t)sel, raw); else memcpy(raw, key.data(), 32); }
    return;
  }
  SplitMix r(mix(seed2 ^ 0xabcdef, i));
  r.fill(key.data(), 32);  // absent w.h.p.
  if (pre) {
    uint8_t slot[32];
    memcpy(slot, key.data(), 32);
    keccak256(slot, 32, key.data());
    if (raw) memcpy(raw, slot, 32);
  } else if (raw) memcpy(raw, key.data(), 32);
}

inline uint32_t pad16(uint32_t n) { return (n + 15u) & ~15u; }

struct Emit {  // one proof after mutation: list of (source bytes, len) + flips + junk
  uint32_t ids[81];
  uint32_t n = 0;
  int flip_node = -1; uint32_t flip_byte = 0; uint8_t flip_mask = 0;
  bool junk = false; uint8_t junk_bytes[48]; uint32_t junk_len = 0;
  bool wrong_key = false;
};

void plan_one(const Trie& t, int64_t sel, uint8_t mut, uint64_t seed2, uint64_t i, Emit& e, H256& key, uint8_t* raw = nullptr) {
  proof_key(t, sel, seed2, i, key, raw);
  Path p;
  walk(t, key, p);
  e.n = p.n;
  memcpy(e.ids, p.ids, sizeof(uint32_t) * p.n);
  SplitMix r(mix(seed2 ^ 0x77, i));
  switch (mut) {
    case MUT_FLIP_LEAF:
      if (e.n) { e.flip_node = (int)e.n - 1; }
      break;
    case MUT_FLIP_INNER:
      if (e.n > 1) e.flip_node = (int)(r.next() % (e.n - 1)); else if (e.n) e.flip_node = 0;
      break;
    case MUT_DROP_LAST: if (e.n) e.n--; break;
    case MUT_DROP_ROOT: if (e.n) { memmove(e.ids, e.ids + 1, sizeof(uint32_t) * (e.n - 1)); e.n--; } break;
    case MUT_WRONG_KEY:
      e.wrong_key = true;
      r.fill(key.data(), 32);
      if (t.raw_keys && t.kind == 1) {
        uint8_t slot[32];
        memcpy(slot, key.data(), 32);
        keccak256(slot, 32, key.data());
        if (raw) memcpy(raw, slot, 32);
      } else if (raw) memcpy(raw, key.data(), 32);
      break;
    case MUT_SHUFFLE:
      for (uint32_t k = e.n; k > 1; k--) { uint32_t j = (uint32_t)(r.next() % k); std::swap(e.ids[k - 1], e.ids[j]); }
      break;
    case MUT_JUNK:
      e.junk = true; e.junk_len = 1 + (uint32_t)(r.next() % 40); r.fill(e.junk_bytes, e.junk_len);
      break;
    default: break;
  }
  if (e.flip_node >= 0) {
    const Node& n = t.p.nodes[e.ids[e.flip_node]];
    e.flip_byte = (uint32_t)(r.next() % n.len);
    e.flip_mask = (uint8_t)(1u << (r.next() & 7));
  }
}

template <class F>
void parallel_for(uint64_t n, int nthreads, F f) {
  if (nthreads <= 1 || n < 1024) { f(0, n); return; }
  std::vector<std::thread> th;
  for (int k = 0; k < nthreads; k++) {
    uint64_t lo = n * k / nthreads, hi = n * (k + 1) / nthreads;
    th.emplace_back([=] { f(lo, hi); });
  }
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

// kind 0: accounts (key = keccak(address), value = account RLP), kind 1: ERC-20 style storage.
// pool_roots (32*P bytes, may be NULL): storage roots assigned to account i as pool[i % P].
void* mptgen_trie_build(uint64_t n_keys, uint64_t seed, int kind, const uint8_t* pool_roots, uint32_t P,
                        int nthreads) {
  if (n_keys > 0xfffffff0ull) return nullptr;
  Trie* t = new Trie();
  t->kind = kind; t->seed = seed;
  for (uint32_t k = 0; k < P; k++) { H256 h; memcpy(h.data(), pool_roots + 32 * k, 32); t->pool_roots.push_back(h); }
  const uint32_t n = (uint32_t)n_keys;
  std::vector<H256> raw(n);
  parallel_for(n, nthreads, [&](uint64_t lo, uint64_t hi) { for (uint64_t i = lo; i < hi; i++) make_key(*t, i, raw[i]); });
  t->key_src.resize(n);
  for (uint32_t i = 0; i < n; i++) t->key_src[i] = i;
  std::sort(t->key_src.begin(), t->key_src.end(), [&](uint32_t a, uint32_t b) { return raw[a] < raw[b]; });
  // drop duplicate keys (cannot happen with keccak outputs, but keep the trie well defined)
  t->keys.resize(n);
  for (uint32_t i = 0; i < n; i++) t->keys[i] = raw[t->key_src[i]];
  t->src_pos.resize(n);
  for (uint32_t i = 0; i < n; i++) t->src_pos[t->key_src[i]] = i;
  std::vector<H256>().swap(raw);
  build_trie(*t, nthreads);
  return t;
}

void mptgen_trie_free(void* h) { delete (Trie*)h; }

void mptgen_trie_info(void* h, uint8_t root[32], uint64_t* n_nodes, uint64_t* arena_bytes, uint64_t* n_keys) {
  Trie* t = (Trie*)h;
  memcpy(root, t->root_hash.data(), 32);
  if (n_nodes) *n_nodes = t->p.nodes.size();
  if (arena_bytes) *arena_bytes = t->p.arena.size();
  if (n_keys) *n_keys = t->keys.size();
}

// key / value of the ORIGINAL index i (value buffer >= 160 bytes); returns the value length
uint32_t mptgen_trie_entry(void* h, uint64_t i, uint8_t key[32], uint8_t* value) {
  Trie* t = (Trie*)h;
  memcpy(key, t->keys[t->src_pos[(uint32_t)i]].data(), 32);
  Buf v;
  make_value(*t, (uint32_t)i, v);
  memcpy(value, v.d, v.n);
  return v.n;
}

// pass 1: node count and padded byte count of each requested proof (after mutation)
int mptgen_proofs_plan(void* h, const int64_t* sel, const uint8_t* mut, uint64_t n, uint64_t seed2,
                       uint32_t* node_count, uint64_t* byte_count, int nthreads) {
  Trie* t = (Trie*)h;
  parallel_for(n, nthreads, [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; i++) {
      Emit e; H256 key;
      plan_one(*t, sel[i], mut ? mut[i] : 0, seed2, i, e, key);
      uint64_t bytes = 0;
      for (uint32_t k = 0; k < e.n; k++) bytes += pad16(t->p.nodes[e.ids[k]].len);
      if (e.junk) bytes += pad16(e.junk_len);
      node_count[i] = e.n + (e.junk ? 1 : 0);
      byte_count[i] = bytes;
    }
  });
  return 0;
}

// pass 2: write proof i into global proof slot slot[i] of a CSR whose prefix sums the caller built
// from pass 1 (proof_first[slot] = first node index, byte_first[slot] = first byte offset).
int mptgen_proofs_emit(void* h, const int64_t* sel, const uint8_t* mut, uint64_t n, uint64_t seed2,
                       const uint64_t* slot, const uint32_t* proof_first, const uint64_t* byte_first,
                       uint8_t* node_bytes, uint64_t* node_off, uint32_t* node_len, uint8_t* roots,
                       uint8_t* keys32, int nthreads, uint8_t* raw32 /* may be NULL: the keys' pre-images (mptgen_trie_raw_keys) */) {
  Trie* t = (Trie*)h;
  parallel_for(n, nthreads, [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; i++) {
      Emit e; H256 key;
      const uint64_t s = slot ? slot[i] : i;
      plan_one(*t, sel[i], mut ? mut[i] : 0, seed2, i, e, key, raw32 ? raw32 + 32 * s : nullptr);
      uint32_t ni = proof_first[s];
      uint64_t bo = byte_first[s];
      for (uint32_t k = 0; k < e.n; k++) {
        const Node& nd = t->p.nodes[e.ids[k]];
        memcpy(node_bytes + bo, t->p.arena.data() + nd.off, nd.len);
        if ((int)k == e.flip_node) node_bytes[bo + e.flip_byte] ^= e.flip_mask;
        node_off[ni] = bo; node_len[ni] = nd.len;
        bo += pad16(nd.len); ni++;
      }
      if (e.junk) {
        memcpy(node_bytes + bo, e.junk_bytes, e.junk_len);
        node_off[ni] = bo; node_len[ni] = e.junk_len;
      }
      memcpy(roots + 32 * s, t->root_hash.data(), 32);
      memcpy(keys32 + 32 * s, key.data(), 32);
    }
  });
  return 0;
}

void mptgen_keccak256(const uint8_t* in, uint64_t len, uint8_t out[32]) { keccak256(in, len, out); }

// storage tries: from now on every proof key has a pre-image (see Trie::raw_keys); set before plan / emit
void mptgen_trie_raw_keys(void* h, int on) { ((Trie*)h)->raw_keys = on != 0; }

// Nested CSR batch -> borsh(StorageProofInput) blobs (crypto-ops/src/types.rs:11-19), one per group: a proof with
// root_from_proof == -1 is an input's account proof (key = address_keccak, root = root_hash), the proofs behind it
// that name it are its storage proofs, raw32 holds their un-hashed storage keys.  account_key is left empty (the
// guest never reads it).  group_first [n_groups + 1] (proof index of every account proof) and blob_off
// [n_groups + 1] are written first (blobs == NULL: returns the total size), then the blobs, multi-threaded.
uint64_t mptgen_csr_to_storage_borsh(const uint8_t* node_bytes, const uint64_t* node_off, const uint32_t* node_len,
                                     const uint32_t* proof_first, uint64_t n_proofs, const uint8_t* roots,
                                     const uint8_t* keys32, const uint8_t* raw32, const int32_t* rfp, uint64_t* group_first,
                                     uint64_t* n_groups_out, uint64_t* blob_off, uint8_t* blobs, int n_threads) {
  uint64_t g = 0;
  for (uint64_t p = 0; p < n_proofs; p++)
    if (rfp[p] < 0) group_first[g++] = p;
  group_first[g] = n_proofs;
  *n_groups_out = g;
  blob_off[0] = 0;
  for (uint64_t i = 0; i < g; i++) {
    const uint64_t a = group_first[i], e = group_first[i + 1];
    uint64_t sz = 4 + 4 + (4 + 32) + 4 + 4 + 32;  // account count, storage count, root_hash, account_key (empty), key count, address_keccak
    for (uint64_t p = a; p < e; p++) {
      if (p > a) sz += 4 + (4 + 32);  // the proof's node count, its key
      for (uint32_t k = proof_first[p]; k < proof_first[p + 1]; k++) sz += 4 + (uint64_t)node_len[k];
    }
    blob_off[i + 1] = blob_off[i] + sz;
  }
  if (!blobs) return blob_off[g];
  auto put32 = [](uint8_t* q, uint32_t v) { q[0] = (uint8_t)v; q[1] = (uint8_t)(v >> 8); q[2] = (uint8_t)(v >> 16); q[3] = (uint8_t)(v >> 24); };
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  std::vector<std::thread> th;
  const uint64_t per = (g + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; t++) {
    const uint64_t lo = std::min(g, per * t), hi = std::min(g, lo + per);
    if (lo >= hi) continue;
    th.emplace_back([=] {
      auto put_proof = [&](uint8_t*& q, uint64_t p) {
        put32(q, proof_first[p + 1] - proof_first[p]); q += 4;
        for (uint32_t k = proof_first[p]; k < proof_first[p + 1]; k++) {
          put32(q, node_len[k]); q += 4;
          memcpy(q, node_bytes + node_off[k], node_len[k]); q += node_len[k];
        }
      };
      for (uint64_t i = lo; i < hi; i++) {
        const uint64_t a = group_first[i], e = group_first[i + 1];
        uint8_t* q = blobs + blob_off[i];
        put_proof(q, a);
        put32(q, (uint32_t)(e - a - 1)); q += 4;
        for (uint64_t p = a + 1; p < e; p++) put_proof(q, p);
        put32(q, 32); q += 4; memcpy(q, roots + 32 * a, 32); q += 32;
        put32(q, 0); q += 4;  // account_key
        put32(q, (uint32_t)(e - a - 1)); q += 4;
        for (uint64_t p = a + 1; p < e; p++) { put32(q, 32); q += 4; memcpy(q, raw32 + 32 * p, 32); q += 32; }
        memcpy(q, keys32 + 32 * a, 32);
      }
    });
  }
  for (auto& x : th) x.join();
  return blob_off[g];
}

// CSR batch -> borsh(MerkleProofInput) blobs, as a prover's input file holds them (crypto-ops/src/types.rs:4-9:
// u32-LE count / length prefixes).  blob_off [n + 1] is written first (call with blobs == NULL to size the
// buffer: returns the total byte count), then the blobs, multi-threaded.
uint64_t mptgen_csr_to_borsh(const uint8_t* node_bytes, const uint64_t* node_off, const uint32_t* node_len,
                             const uint32_t* proof_first, uint64_t n_proofs, const uint8_t* roots,
                             const uint8_t* key_bytes, const uint32_t* key_off, uint64_t* blob_off, uint8_t* blobs,
                             int n_threads) {
  blob_off[0] = 0;
  for (uint64_t p = 0; p < n_proofs; p++) {
    uint64_t sz = 4 + 4 + 32 + 4 + (uint64_t)(key_off[p + 1] - key_off[p]);
    for (uint32_t i = proof_first[p]; i < proof_first[p + 1]; i++) sz += 4 + (uint64_t)node_len[i];
    blob_off[p + 1] = blob_off[p] + sz;
  }
  if (!blobs) return blob_off[n_proofs];
  auto put32 = [](uint8_t* q, uint32_t v) { q[0] = (uint8_t)v; q[1] = (uint8_t)(v >> 8); q[2] = (uint8_t)(v >> 16); q[3] = (uint8_t)(v >> 24); };
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  std::vector<std::thread> th;
  const uint64_t per = (n_proofs + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; t++) {
    const uint64_t lo = std::min(n_proofs, per * t), hi = std::min(n_proofs, lo + per);
    if (lo >= hi) continue;
    th.emplace_back([=] {
      for (uint64_t p = lo; p < hi; p++) {
        uint8_t* q = blobs + blob_off[p];
        put32(q, proof_first[p + 1] - proof_first[p]); q += 4;
        for (uint32_t i = proof_first[p]; i < proof_first[p + 1]; i++) {
          put32(q, node_len[i]); q += 4;
          memcpy(q, node_bytes + node_off[i], node_len[i]); q += node_len[i];
        }
        put32(q, 32); q += 4;
        memcpy(q, roots + 32 * p, 32); q += 32;
        const uint32_t kl = key_off[p + 1] - key_off[p];
        put32(q, kl); q += 4;
        if (kl) memcpy(q, key_bytes + key_off[p], kl);
      }
    });
  }
  for (auto& x : th) x.join();
  return blob_off[n_proofs];
}

}  // extern "C"
