/*
 * oracle/rv32emu.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A small RV32IM interpreter with the six SP1 ecalls the reference's committed
 * guest binary uses.  It executes the reference's OWN compiled
 * `verify_merkle_proof` (crypto-ops/src/lib.rs:8-23 + eth_trie@ade617b +
 * tiny-keccak 2.0.2, built from circuits/sp1-merkle-proof/src/main.rs:1-14)
 * from /root/reference/circuits/elf/riscv32im-succinct-zkvm-elf, bit-exactly.
 * The ELF is read at run time from the path the caller passes; nothing of the
 * reference is copied into this repository.
 *
 * Used only (a) in this container to generate tests/golden/ fixtures
 * (oracle/gen_golden.py) and (b) by `-m "not gpu"` tests when /root/reference
 * is mounted, to fuzz the C restatement in oracle/mpt_oracle.c against the
 * reference itself.  Never shipped, never timed as a performance baseline.
 *
 * Interface follows SURVEY.md Appendix B:
 *   hint 0 = u64_le(len(p)) || p ,  p = borsh(MerkleProofInput)
 *   ecall t0: 0xF0 HINT_LEN (-> t0... see below), 0xF1 HINT_READ(a0=ptr,a1=len),
 *             2 WRITE(a0=fd,a1=ptr,a2=len), 0x10 COMMIT, 0x1A COMMIT_DEFERRED,
 *             0 HALT(a0=exit code)
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <sys/mman.h>

#define MEM_SIZE (1ull << 31) /* SP1 guest addresses are < 0x78000000 */

typedef struct {
  uint8_t *mem;
  uint32_t entry;
  /* image kept for fast reset */
  struct { uint32_t vaddr, filesz, memsz; const uint8_t *src; } seg[8];
  int nseg;
  uint8_t *elf_copy;
  uint32_t hi_water; /* highest heap address touched (for reset) */
} rv32_vm;

static uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

rv32_vm *rv32_load(const uint8_t *elf, uint64_t elf_len) {
  if (elf_len < 52 || memcmp(elf, "\x7f" "ELF", 4) != 0 || elf[4] != 1) return NULL;
  if (rd16(elf + 18) != 243) return NULL; /* EM_RISCV */
  rv32_vm *vm = (rv32_vm *)calloc(1, sizeof(rv32_vm));
  vm->elf_copy = (uint8_t *)malloc(elf_len);
  memcpy(vm->elf_copy, elf, elf_len);
  vm->entry = rd32(elf + 24);
  uint32_t phoff = rd32(elf + 28);
  uint16_t phentsize = rd16(elf + 42), phnum = rd16(elf + 44);
  for (int i = 0; i < phnum && vm->nseg < 8; i++) {
    const uint8_t *ph = vm->elf_copy + phoff + (uint32_t)i * phentsize;
    if (rd32(ph) != 1) continue; /* PT_LOAD */
    vm->seg[vm->nseg].src = vm->elf_copy + rd32(ph + 4);
    vm->seg[vm->nseg].vaddr = rd32(ph + 8);
    vm->seg[vm->nseg].filesz = rd32(ph + 16);
    vm->seg[vm->nseg].memsz = rd32(ph + 20);
    vm->nseg++;
  }
  vm->mem = (uint8_t *)mmap(NULL, MEM_SIZE, PROT_READ | PROT_WRITE,
                            MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (vm->mem == MAP_FAILED) { free(vm->elf_copy); free(vm); return NULL; }
  return vm;
}

void rv32_free(rv32_vm *vm) {
  if (!vm) return;
  munmap(vm->mem, MEM_SIZE);
  free(vm->elf_copy);
  free(vm);
}

static void vm_reset(rv32_vm *vm) {
  /* drop every page the previous run dirtied; they read back as zero */
  madvise(vm->mem, MEM_SIZE, MADV_DONTNEED);
  for (int i = 0; i < vm->nseg; i++)
    memcpy(vm->mem + vm->seg[i].vaddr, vm->seg[i].src, vm->seg[i].filesz);
}

typedef struct {
  int32_t exit_code;      /* guest exit code; -1 = emulator fault; -2 = step limit */
  uint64_t steps;         /* instructions fetched incl. the halting ecall */
  uint32_t pub_len;       /* bytes written to fd 3 (public values) */
  uint32_t err_len;       /* bytes written to fd 2 (stderr) */
  uint32_t fault_pc;
  uint32_t fault_insn;
} rv32_result;

#define MASK31 0x7fffffffu

/* Run the loaded guest once with a single hint buffer. */
int rv32_run(rv32_vm *vm, const uint8_t *hint, uint32_t hint_len,
             uint8_t *pub, uint32_t pub_cap, uint8_t *err, uint32_t err_cap,
             uint64_t max_steps, rv32_result *res) {
  vm_reset(vm);
  uint8_t *M = vm->mem;
  uint32_t x[32];
  memset(x, 0, sizeof x);
  uint32_t pc = vm->entry;
  uint64_t steps = 0;
  int hint_taken = 0;
  res->pub_len = res->err_len = 0;
  res->exit_code = -1;
  res->fault_pc = res->fault_insn = 0;
  if (max_steps == 0) max_steps = 1ull << 34;

  for (;;) {
    if (steps >= max_steps) { res->exit_code = -2; break; }
    uint32_t insn = rd32(M + (pc & MASK31));
    steps++;
    uint32_t opc = insn & 0x7f, rd = (insn >> 7) & 31, f3 = (insn >> 12) & 7;
    uint32_t rs1 = (insn >> 15) & 31, rs2 = (insn >> 20) & 31, f7 = insn >> 25;
    uint32_t a = x[rs1], b = x[rs2], npc = pc + 4, v = 0;
    int wr = 1;
    switch (opc) {
    case 0x37: v = insn & 0xfffff000u; break;                    /* LUI */
    case 0x17: v = pc + (insn & 0xfffff000u); break;             /* AUIPC */
    case 0x6f: {                                                 /* JAL */
      int32_t imm = (int32_t)(((insn >> 31) ? 0xfff00000u : 0) | (insn & 0xff000) |
                              ((insn >> 9) & 0x800) | ((insn >> 20) & 0x7fe));
      v = pc + 4; npc = pc + (uint32_t)imm; break; }
    case 0x67: v = pc + 4; npc = (a + (uint32_t)((int32_t)insn >> 20)) & ~1u; break; /* JALR */
    case 0x63: {                                                 /* BRANCH */
      int32_t imm = (int32_t)(((insn >> 31) ? 0xfffff000u : 0) | ((insn << 4) & 0x800) |
                              ((insn >> 20) & 0x7e0) | ((insn >> 7) & 0x1e));
      int t;
      switch (f3) {
      case 0: t = a == b; break;
      case 1: t = a != b; break;
      case 4: t = (int32_t)a < (int32_t)b; break;
      case 5: t = (int32_t)a >= (int32_t)b; break;
      case 6: t = a < b; break;
      case 7: t = a >= b; break;
      default: goto fault;
      }
      if (t) npc = pc + (uint32_t)imm;
      wr = 0; break; }
    case 0x03: {                                                 /* LOAD */
      uint32_t ad = (a + (uint32_t)((int32_t)insn >> 20)) & MASK31;
      switch (f3) {
      case 0: v = (uint32_t)(int32_t)(int8_t)M[ad]; break;
      case 1: v = (uint32_t)(int32_t)(int16_t)rd16(M + ad); break;
      case 2: v = rd32(M + ad); break;
      case 4: v = M[ad]; break;
      case 5: v = rd16(M + ad); break;
      default: goto fault;
      }
      break; }
    case 0x23: {                                                 /* STORE */
      int32_t imm = (int32_t)(((int32_t)insn >> 25) << 5) | (int32_t)((insn >> 7) & 31);
      uint32_t ad = (a + (uint32_t)imm) & MASK31;
      switch (f3) {
      case 0: M[ad] = (uint8_t)b; break;
      case 1: { uint16_t h = (uint16_t)b; memcpy(M + ad, &h, 2); break; }
      case 2: memcpy(M + ad, &b, 4); break;
      default: goto fault;
      }
      wr = 0; break; }
    case 0x13: {                                                 /* OP-IMM */
      uint32_t imm = (uint32_t)((int32_t)insn >> 20), sh = rs2;
      switch (f3) {
      case 0: v = a + imm; break;
      case 1: v = a << sh; break;
      case 2: v = (int32_t)a < (int32_t)imm; break;
      case 3: v = a < imm; break;
      case 4: v = a ^ imm; break;
      case 5: v = (f7 & 0x20) ? (uint32_t)((int32_t)a >> sh) : a >> sh; break;
      case 6: v = a | imm; break;
      case 7: v = a & imm; break;
      }
      break; }
    case 0x33: {                                                 /* OP */
      if (f7 == 1) {                                             /* M extension */
        switch (f3) {
        case 0: v = a * b; break;
        case 1: v = (uint32_t)(((int64_t)(int32_t)a * (int64_t)(int32_t)b) >> 32); break;
        case 2: v = (uint32_t)(((int64_t)(int32_t)a * (int64_t)(uint64_t)b) >> 32); break;
        case 3: v = (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); break;
        case 4: v = b == 0 ? 0xffffffffu
                  : (a == 0x80000000u && b == 0xffffffffu) ? a
                  : (uint32_t)((int32_t)a / (int32_t)b); break;
        case 5: v = b == 0 ? 0xffffffffu : a / b; break;
        case 6: v = b == 0 ? a
                  : (a == 0x80000000u && b == 0xffffffffu) ? 0
                  : (uint32_t)((int32_t)a % (int32_t)b); break;
        case 7: v = b == 0 ? a : a % b; break;
        }
      } else {
        switch (f3) {
        case 0: v = (f7 & 0x20) ? a - b : a + b; break;
        case 1: v = a << (b & 31); break;
        case 2: v = (int32_t)a < (int32_t)b; break;
        case 3: v = a < b; break;
        case 4: v = a ^ b; break;
        case 5: v = (f7 & 0x20) ? (uint32_t)((int32_t)a >> (b & 31)) : a >> (b & 31); break;
        case 6: v = a | b; break;
        case 7: v = a & b; break;
        }
      }
      break; }
    case 0x0f: wr = 0; break;                                    /* FENCE */
    case 0x73: {                                                 /* SYSTEM */
      wr = 0;
      if (insn != 0x00000073u) goto fault;                       /* only ECALL */
      uint32_t code = x[5];
      switch (code) {
      case 0x00: res->exit_code = (int32_t)x[10]; goto done;     /* HALT */
      case 0xF0: x[5] = hint_taken ? 0 : hint_len; break;        /* HINT_LEN -> t0 */
      case 0xF1: {                                               /* HINT_READ */
        uint32_t p = x[10] & MASK31, n = x[11];
        if (hint_taken || n > hint_len) goto fault;
        memcpy(M + p, hint, n);
        hint_taken = 1;
        break; }
      case 0x02: {                                               /* WRITE */
        uint32_t fd = x[10], p = x[11] & MASK31, n = x[12];
        if (fd == 3) {
          uint32_t k = n; if (res->pub_len + k > pub_cap) k = pub_cap - res->pub_len;
          memcpy(pub + res->pub_len, M + p, k); res->pub_len += k;
        } else if (fd == 2) {
          uint32_t k = n; if (res->err_len + k > err_cap) k = err_cap - res->err_len;
          memcpy(err + res->err_len, M + p, k); res->err_len += k;
        }
        break; }
      case 0x10: case 0x1A: break;                     /* COMMIT* */
      default: goto fault;
      }
      break; }
    default: goto fault;
    }
    if (wr && rd) x[rd] = v;
    pc = npc;
    continue;
  fault:
    res->fault_pc = pc; res->fault_insn = insn; res->exit_code = -1;
    break;
  }
done:
  res->steps = steps;
  return res->exit_code;
}
