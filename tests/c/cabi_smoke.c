/* plain-C consumer of include/mptv.h: the header must compile as C99 and the library must link without C++ */
#include "mptv.h"
#include <stdio.h>
#include <string.h>

static void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

int main(void) {
  mptv_ctx* ctx = NULL;
  int rc = mptv_create(NULL, 0, &ctx);
  int fail = 0;
  printf("mptv_create -> %d (%s)\n", rc, mptv_strerror(rc));
  uint8_t key[9];
  printf("rlp(300) has %u bytes\n", mptv_rlp_index(300, key));

  /* the storage guest's Account decode on the host: rlp([5, 0x0100, 0x11 x 32, 0x22 x 32]) */
  uint8_t acct[72], root[32];
  acct[0] = 0xf8; acct[1] = 70; acct[2] = 0x05; acct[3] = 0x82; acct[4] = 0x01; acct[5] = 0x00;
  acct[6] = 0xa0; memset(acct + 7, 0x11, 32); acct[39] = 0xa0; memset(acct + 40, 0x22, 32);
  if (mptv_account_storage_root(acct, 72, root) != 1 || root[0] != 0x11 || mptv_account_storage_root(acct, 71, NULL) != 0) fail |= 1;

  /* the smallest borsh(StorageProofInput): no account proof, no storage proofs, a 32-byte root, no keys, address_keccak */
  uint8_t blob[84 + 16];
  uint64_t off[2] = {0, 84}, first[2] = {7, 7};
  memset(blob, 0, sizeof blob);
  put32(blob, 0); put32(blob + 4, 0); put32(blob + 8, 32); memset(blob + 12, 0xab, 32);
  put32(blob + 44, 0); put32(blob + 48, 0); memset(blob + 52, 0xcd, 32);
  mptv_host_batch* hb = NULL;
  const uint8_t* hk = NULL;
  rc = mptv_flatten_storage_borsh(blob, off, 1, 1, 0, 0, &hb, NULL, first, &hk);
  if (rc != MPTV_OK || first[0] != 0 || first[1] != 1 || mptv_host_batch_view(hb)->n_proofs != 1 || hk[0] != 0) fail |= 2;
  if (hb) mptv_host_batch_free(hb);
  off[1] = 83;  /* one byte short: borsh::from_slice would fail */
  hb = NULL;
  if (mptv_flatten_storage_borsh(blob, off, 1, 1, 0, 0, &hb, NULL, first, &hk) != MPTV_ERR_ARG) fail |= 4;
  off[1] = 84;

  if (ctx) {  /* with a B200: the guest's outcome for that input -- an empty proof cannot hash to the root */
    uint8_t status[1], ist[1];
    uint64_t voff[1];
    uint32_t vlen[1];
    mptv_result res;
    res.status = status; res.value_off = voff; res.value_len = vlen;
    rc = mptv_verify_storage_borsh(ctx, blob, off, 1, 0, first, ist, 1, &res);
    printf("mptv_verify_storage_borsh -> %d, input status %d\n", rc, (int)ist[0]);
    if (rc != MPTV_OK || ist[0] != MPTV_ST_INVALID_STATE_ROOT || first[1] != 1) fail |= 8;
    mptv_destroy(ctx);
  }
  printf("storage entries: %s\n", fail ? "FAILED" : "ok");
  return fail;
}
