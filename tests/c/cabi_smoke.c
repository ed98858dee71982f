/* plain-C consumer of include/mptv.h: the header must compile as C99 and the library must link without C++ */
#include "mptv.h"
#include <stdio.h>
int main(void) {
  mptv_ctx* ctx = NULL;
  int rc = mptv_create(NULL, 0, &ctx);
  printf("mptv_create -> %d (%s)\n", rc, mptv_strerror(rc));
  uint8_t key[9];
  printf("rlp(300) has %u bytes\n", mptv_rlp_index(300, key));
  if (ctx) mptv_destroy(ctx);
  return 0;
}
