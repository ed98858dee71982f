"""Issue-rate probes of the integer pipes (alu: LOP3/SHF, fma: IMAD/IMAD.HI/IMAD.WIDE) and mixes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_state_proofs_b200 as z
v = z.Verifier([0])
names = ["LOP3", "SHF", "keccak mix 11 LOP3:5 SHF", "IMAD.lo", "IMAD.HI", "IMAD.WIDE", "LOP3+IMAD.lo 1:1",
         "LOP3+IMAD.HI 1:1", "LOP3+IMAD.WIDE 2:1"]
for m, n in enumerate(names):
    print(f"mode {m} {n:28s} {v.int_issue_peak(0, m) / 1e12:7.2f} T lane-instr/s")
