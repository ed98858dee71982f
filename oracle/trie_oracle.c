/*
 * oracle/trie_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the transaction / receipt trie REBUILD that trie-utils performs with
 * eth_trie 0.5.0 (crates.io, Cargo.lock:2794-2806):
 *     trie = EthTrie::new(MemoryDB); for (i, tx): trie.insert(rlp(i), tx.eip2718_encode());
 *     trie.root_hash(); trie.get_proof(rlp(target))
 *   /root/reference/trie-utils/src/proofs/transaction.rs:41-68, :83-119
 *   /root/reference/trie-utils/src/proofs/receipt.rs:49-86, /root/reference/trie-utils/src/receipt.rs:8-38
 *
 * eth_trie's source is not under /root/reference and -- unlike verify_merkle_proof -- insert /
 * get_proof are not in the committed guest ELF, so this follows the published algorithm (Ethereum
 * Yellow Paper appendix D, the same sequential node surgery eth_trie::insert_at performs: leaf
 * split, extension split, branch descent; commit = bottom-up encode, children >= 32 bytes by hash).
 * PARITY PIN: the roots and proofs produced here are fed to the reference's own
 * verify_merkle_proof (the guest ELF under oracle/rv32emu.c) in tests/test_rebuild_oracle.py and in
 * tests/golden/rebuild_vectors.json.gz: the reference accepts every proof against these roots and
 * returns the inserted bytes; its lib.rs:19 re-encode assert independently confirms the root
 * node's canonical encoding.  A second, independently written builder (oracle/pytrie.py, sorted
 * recursion instead of sequential insertion) must give identical roots.
 *
 * insert(key, b"") is a delete in eth_trie; items are applied in order (last write wins), so the
 * final map is computed first and only live entries are inserted -- the MPT of a key/value set is
 * unique, so the result is the same trie.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

uint32_t mpto_keccak256(const uint8_t *in, uint64_t len, uint8_t out[32]);

typedef struct tnode {
  int kind;                 /* 0 leaf, 1 ext, 2 branch */
  uint8_t *path; uint32_t plen;         /* nibbles (leaf: remaining key, ext: shared prefix) */
  const uint8_t *val; uint32_t vlen;    /* leaf value / branch value (vlen 0 = none) */
  struct tnode *child[16];              /* branch children; ext child in child[0] */
  uint8_t *enc; uint32_t elen;          /* encoding (after commit) */
  uint8_t hash[32];
} tnode;

typedef struct tpool { uint8_t **blocks; size_t nblocks, cap, used; } tpool;

static void *palloc(tpool *P, size_t n) {
  n = (n + 15) & ~(size_t)15;
  if (P->nblocks == 0 || P->used + n > P->cap) {
    size_t c = n > (1u << 20) ? n : (1u << 20);
    P->blocks = (uint8_t **)realloc(P->blocks, sizeof(uint8_t *) * (P->nblocks + 1));
    P->blocks[P->nblocks++] = (uint8_t *)malloc(c);
    P->cap = c; P->used = 0;
  }
  void *p = P->blocks[P->nblocks - 1] + P->used;
  P->used += n;
  memset(p, 0, n);
  return p;
}
static void pfree(tpool *P) {
  for (size_t i = 0; i < P->nblocks; i++) free(P->blocks[i]);
  free(P->blocks);
  memset(P, 0, sizeof *P);
}

static tnode *new_leaf(tpool *P, const uint8_t *path, uint32_t plen, const uint8_t *v, uint32_t vl) {
  tnode *n = (tnode *)palloc(P, sizeof(tnode));
  n->kind = 0; n->path = (uint8_t *)palloc(P, plen + 1); memcpy(n->path, path, plen); n->plen = plen;
  n->val = v; n->vlen = vl;
  return n;
}

/* EthTrie::insert_at */
static tnode *insert(tpool *P, tnode *n, const uint8_t *path, uint32_t plen, const uint8_t *v, uint32_t vl) {
  if (!n) return new_leaf(P, path, plen, v, vl);
  if (n->kind == 2) {
    if (plen == 0) { n->val = v; n->vlen = vl; return n; }
    n->child[path[0]] = insert(P, n->child[path[0]], path + 1, plen - 1, v, vl);
    return n;
  }
  uint32_t m = 0;
  while (m < n->plen && m < plen && n->path[m] == path[m]) m++;
  if (n->kind == 0) {
    if (m == n->plen && m == plen) { n->val = v; n->vlen = vl; return n; }
    tnode *b = (tnode *)palloc(P, sizeof(tnode));
    b->kind = 2;
    if (m == n->plen) { b->val = n->val; b->vlen = n->vlen; }
    else b->child[n->path[m]] = new_leaf(P, n->path + m + 1, n->plen - m - 1, n->val, n->vlen);
    if (m == plen) { b->val = v; b->vlen = vl; }
    else b->child[path[m]] = new_leaf(P, path + m + 1, plen - m - 1, v, vl);
    if (m == 0) return b;
    tnode *e = (tnode *)palloc(P, sizeof(tnode));
    e->kind = 1; e->path = (uint8_t *)palloc(P, m); memcpy(e->path, path, m); e->plen = m; e->child[0] = b;
    return e;
  }
  /* extension */
  if (m == n->plen) { n->child[0] = insert(P, n->child[0], path + m, plen - m, v, vl); return n; }
  tnode *b = (tnode *)palloc(P, sizeof(tnode));
  b->kind = 2;
  /* the rest of the old extension hangs under its first differing nibble */
  if (n->plen - m - 1 == 0) b->child[n->path[m]] = n->child[0];
  else {
    tnode *e2 = (tnode *)palloc(P, sizeof(tnode));
    e2->kind = 1; e2->plen = n->plen - m - 1; e2->path = (uint8_t *)palloc(P, e2->plen);
    memcpy(e2->path, n->path + m + 1, e2->plen); e2->child[0] = n->child[0];
    b->child[n->path[m]] = e2;
  }
  if (m == plen) { b->val = v; b->vlen = vl; }
  else b->child[path[m]] = new_leaf(P, path + m + 1, plen - m - 1, v, vl);
  if (m == 0) return b;
  tnode *e = (tnode *)palloc(P, sizeof(tnode));
  e->kind = 1; e->path = (uint8_t *)palloc(P, m); memcpy(e->path, path, m); e->plen = m; e->child[0] = b;
  return e;
}

/* ---- encoding */
static uint32_t put_hdr(uint8_t *o, uint32_t n, int list) {
  uint8_t base = list ? 0xC0 : 0x80;
  if (n < 56) { o[0] = (uint8_t)(base + n); return 1; }
  uint8_t t[4]; int k = 0;
  for (uint32_t v = n; v; v >>= 8) t[k++] = (uint8_t)v;
  o[0] = (uint8_t)(base + 55 + k);
  for (int i = 0; i < k; i++) o[1 + i] = t[k - 1 - i];
  return 1 + (uint32_t)k;
}

static uint32_t put_str(uint8_t *o, const uint8_t *s, uint32_t n) {
  if (n == 1 && s[0] < 0x80) { o[0] = s[0]; return 1; }
  uint32_t h = put_hdr(o, n, 0);
  memcpy(o + h, s, n);
  return h + n;
}
static uint32_t put_hp(uint8_t *o, const uint8_t *nb, uint32_t n, int leaf) {
  uint8_t t[80];
  uint32_t k = 0, i = 0;
  uint8_t flag = leaf ? 0x20 : 0x00;
  if (n & 1) { t[k++] = (uint8_t)(flag | 0x10 | nb[0]); i = 1; } else t[k++] = flag;
  for (; i < n; i += 2) t[k++] = (uint8_t)((nb[i] << 4) | nb[i + 1]);
  return put_str(o, t, k);
}

typedef struct { uint64_t perms; uint64_t nodes_hashed; } tstats;

static void commit(tpool *P, tnode *n, int is_root, tstats *st) {
  uint32_t cap = 0;
  if (n->kind == 0) cap = 16 + n->plen / 2 + 8 + n->vlen;
  else if (n->kind == 1) { commit(P, n->child[0], 0, st); cap = 16 + n->plen / 2 + 40; }
  else {
    cap = 16 + 8 + n->vlen;
    for (int i = 0; i < 16; i++) if (n->child[i]) { commit(P, n->child[i], 0, st); cap += 33; } else cap += 1;
  }
  uint8_t *pl = (uint8_t *)palloc(P, cap + 8);
  uint32_t k = 0;
  if (n->kind == 0) { k += put_hp(pl + k, n->path, n->plen, 1); k += put_str(pl + k, n->val, n->vlen); }
  else if (n->kind == 1) {
    k += put_hp(pl + k, n->path, n->plen, 0);
    tnode *c = n->child[0];
    if (c->elen < 32) { memcpy(pl + k, c->enc, c->elen); k += c->elen; } else k += put_str(pl + k, c->hash, 32);
  } else {
    for (int i = 0; i < 16; i++) {
      tnode *c = n->child[i];
      if (!c) pl[k++] = 0x80;
      else if (c->elen < 32) { memcpy(pl + k, c->enc, c->elen); k += c->elen; }
      else k += put_str(pl + k, c->hash, 32);
    }
    if (n->vlen) k += put_str(pl + k, n->val, n->vlen); else pl[k++] = 0x80;
  }
  n->enc = (uint8_t *)palloc(P, k + 8);
  uint32_t h = put_hdr(n->enc, k, 1);
  memcpy(n->enc + h, pl, k);
  n->elen = h + k;
  if (n->elen >= 32 || is_root) { st->perms += mpto_keccak256(n->enc, n->elen, n->hash); st->nodes_hashed++; }
}

static const uint8_t EMPTY_ROOT[32] = {0x56, 0xe8, 0x1f, 0x17, 0x1b, 0xcc, 0x55, 0xa6, 0xff, 0x83, 0x45, 0xe6, 0x92, 0xc0, 0xf8, 0x6e,
                                       0x5b, 0x48, 0xe0, 0x1b, 0x99, 0x6c, 0xad, 0xc0, 0x01, 0x62, 0x2f, 0xb5, 0xe3, 0x63, 0xb4, 0x21};

typedef struct {
  tpool pool;
  tnode *root;
  tstats st;
} ttrie;

static __thread const uint8_t *g_kb;
static __thread const uint32_t *g_ko;
static int cmp_item(const void *x, const void *y) {
  uint64_t a = *(const uint64_t *)x, b = *(const uint64_t *)y;
  uint32_t la = g_ko[a + 1] - g_ko[a], lb = g_ko[b + 1] - g_ko[b];
  int c = memcmp(g_kb + g_ko[a], g_kb + g_ko[b], la < lb ? la : lb);
  if (c) return c;
  if (la != lb) return la < lb ? -1 : 1;
  return a < b ? -1 : (a > b ? 1 : 0);
}

/* build from n items in insertion order (last write wins, empty value deletes) */
static void build(ttrie *T, const uint8_t *key_bytes, const uint32_t *key_off, const uint8_t *val_bytes,
                  const uint64_t *val_off, const uint32_t *val_len, uint64_t first, uint64_t n) {
  memset(T, 0, sizeof *T);
  /* final map: an item is live iff no LATER item has the same key (sort by key, then index) */
  uint64_t *idx = (uint64_t *)malloc(sizeof(uint64_t) * (n ? n : 1));
  uint8_t *live = (uint8_t *)malloc(n ? n : 1);
  for (uint64_t i = 0; i < n; i++) { idx[i] = i; live[i] = 1; }
  g_kb = key_bytes; g_ko = key_off + first;
  qsort(idx, n, sizeof(uint64_t), cmp_item);
  for (uint64_t i = 0; i + 1 < n; i++) {
    uint64_t a = idx[i], b = idx[i + 1];
    uint32_t la = g_ko[a + 1] - g_ko[a], lb = g_ko[b + 1] - g_ko[b];
    if (la == lb && memcmp(g_kb + g_ko[a], g_kb + g_ko[b], la) == 0) live[a] = 0; /* b > a comes later */
  }
  uint8_t nib[2 * 64 + 2];
  for (uint64_t i = 0; i < n; i++) {
    uint64_t it = first + i;
    if (!live[i]) continue;
    uint32_t kl = key_off[it + 1] - key_off[it];
    if (kl > 64) kl = 64;
    const uint8_t *k = key_bytes + key_off[it];
    for (uint32_t b = 0; b < kl; b++) { nib[2 * b] = k[b] >> 4; nib[2 * b + 1] = k[b] & 15; }
    if (val_len[it] == 0) continue; /* delete of a key that (being last) is therefore absent */
    T->root = insert(&T->pool, T->root, nib, 2 * kl, val_bytes + val_off[it], val_len[it]);
  }
  free(idx); free(live);
  if (T->root) commit(&T->pool, T->root, 1, &T->st);
}

/* roots32[32t..] = MPT root of trie t = items [trie_first[t], trie_first[t+1]) */
typedef struct {
  const uint8_t *key_bytes; const uint32_t *key_off; const uint8_t *val_bytes; const uint64_t *val_off;
  const uint32_t *val_len; const uint32_t *trie_first; uint64_t lo, hi; uint8_t *roots; uint64_t perms, nodes;
} rjob;

static void *rworker(void *arg) {
  rjob *j = (rjob *)arg;
  for (uint64_t t = j->lo; t < j->hi; t++) {
    ttrie T;
    build(&T, j->key_bytes, j->key_off, j->val_bytes, j->val_off, j->val_len, j->trie_first[t],
          j->trie_first[t + 1] - j->trie_first[t]);
    memcpy(j->roots + 32 * t, T.root ? T.root->hash : EMPTY_ROOT, 32);
    j->perms += T.st.perms; j->nodes += T.st.nodes_hashed;
    pfree(&T.pool);
  }
  return NULL;
}

int mpto_trie_roots(const uint8_t *key_bytes, const uint32_t *key_off, const uint8_t *val_bytes,
                    const uint64_t *val_off, const uint32_t *val_len, const uint32_t *trie_first, uint64_t n_tries,
                    uint8_t *roots32, int nthreads, uint64_t *perms, uint64_t *nodes_hashed) {
  if (nthreads < 1) nthreads = 1;
  if ((uint64_t)nthreads > n_tries) nthreads = n_tries ? (int)n_tries : 1;
  rjob *jobs = (rjob *)calloc((size_t)nthreads, sizeof(rjob));
  pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; t++) {
    rjob *j = &jobs[t];
    j->key_bytes = key_bytes; j->key_off = key_off; j->val_bytes = val_bytes; j->val_off = val_off;
    j->val_len = val_len; j->trie_first = trie_first; j->roots = roots32;
    j->lo = n_tries * (uint64_t)t / (uint64_t)nthreads; j->hi = n_tries * (uint64_t)(t + 1) / (uint64_t)nthreads;
    pthread_create(&th[t], NULL, rworker, j);
  }
  uint64_t p = 0, nh = 0;
  for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); p += jobs[t].perms; nh += jobs[t].nodes; }
  if (perms) *perms = p;
  if (nodes_hashed) *nodes_hashed = nh;
  free(jobs); free(th);
  return 0;
}

/* get_proof for ONE trie: encoded nodes on the path of `key`, root first; the root always, other
 * nodes only when referenced by hash.  Output: concatenated nodes + lengths.  Returns node count. */
int mpto_trie_get_proof(const uint8_t *key_bytes, const uint32_t *key_off, const uint8_t *val_bytes,
                        const uint64_t *val_off, const uint32_t *val_len, uint64_t first, uint64_t n,
                        const uint8_t *key, uint32_t key_len, uint8_t *out, uint64_t out_cap, uint32_t *out_lens,
                        uint32_t max_nodes, uint8_t root[32]) {
  ttrie T;
  build(&T, key_bytes, key_off, val_bytes, val_off, val_len, first, n);
  memcpy(root, T.root ? T.root->hash : EMPTY_ROOT, 32);
  int cnt = 0;
  uint64_t used = 0;
  uint8_t nib[2 * 64 + 2];
  if (key_len > 64) key_len = 64;
  for (uint32_t b = 0; b < key_len; b++) { nib[2 * b] = key[b] >> 4; nib[2 * b + 1] = key[b] & 15; }
  uint32_t plen = 2 * key_len, idx = 0;
  tnode *c = T.root;
  while (c) {
    if ((c->elen >= 32 || c == T.root) && (uint32_t)cnt < max_nodes && used + c->elen <= out_cap) {
      memcpy(out + used, c->enc, c->elen); out_lens[cnt++] = c->elen; used += c->elen;
    }
    if (c->kind == 0) break;
    if (c->kind == 1) {
      if (plen - idx < c->plen || memcmp(c->path, nib + idx, c->plen) != 0) break;
      idx += c->plen; c = c->child[0];
    } else {
      if (idx >= plen) break;
      c = c->child[nib[idx]]; idx++;
    }
  }
  pfree(&T.pool);
  return cnt;
}
