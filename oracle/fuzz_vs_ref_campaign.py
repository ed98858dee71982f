"""oracle/fuzz_vs_ref_campaign.py -- TEST INFRASTRUCTURE.  A timed campaign of oracle/fuzz_vs_ref.py's comparison
(C restatement, plain and mirror mode, vs the reference's guest ELF under oracle/rv32emu.c) over consecutive seeds,
with the deep-nesting family (inline nodes 64 ... 2 000 levels deep, planted decode faults) added to every seed.
    python -m oracle.fuzz_vs_ref_campaign [minutes] [first_seed]      (this container: needs /root/reference)"""
import random
import sys
import time
from collections import Counter
from concurrent.futures import ThreadPoolExecutor

from .fuzzgen import corpus, deep_nested_cases
from .pyoracle import Oracle, RefElf


def main():
    minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12000
    o, ref = Oracle(), RefElf()
    t0 = time.time()
    total = bad = 0
    hist = Counter()

    def one(c):
        a = o.verify(c["root"], c["proof"], c["key"])
        a2 = o.verify(c["root"], c["proof"], c["key"], mirror=True)
        b = ref.run(c["root"], c["proof"], c["key"])
        return a, a2, b

    while time.time() - t0 < 60 * minutes:
        ts = time.time()
        cases = corpus(seed, o.keccak256, 300, 12000, 12000, 5000, 12000, 3000)
        cases += deep_nested_cases(random.Random(seed), o.keccak256, 27)
        with ThreadPoolExecutor(8) as ex:
            res = list(ex.map(one, cases))
        nb = 0
        for c, (a, a2, b) in zip(cases, res):
            hist[b["status"]] += 1
            if not ((a[0] == b["status"]) and (a[1] == b["value"]) and (a2[0] == a[0]) and (a2[1] == a[1])):
                nb += 1
                if bad + nb <= 10:
                    print("MISMATCH", seed, c["tag"], "oracle:", a[0], "ref:", b["status"], flush=True)
        bad += nb
        total += len(cases)
        print(f"seed {seed}: {len(cases)} cases, {nb} mismatches, {time.time() - ts:.1f}s", flush=True)
        seed += 1
    print(f"TOTAL {total} cases: {bad} mismatches in verdict class or value bytes; reference verdicts {dict(sorted(hist.items()))}; "
          f"{(time.time() - t0) / 60:.1f} min")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
