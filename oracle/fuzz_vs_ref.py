"""oracle/fuzz_vs_ref.py -- TEST INFRASTRUCTURE.  Live differential fuzz of the C restatement
(oracle/mpt_oracle.c) against the reference's own guest ELF.  Usage:
    python -m oracle.fuzz_vs_ref [seed] [n_tries] [n_mut] [n_weird] [n_nested] [n_reseal] [n_padded]
"""
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from collections import Counter

from .pyoracle import Oracle, RefElf, STATUS_NAMES
from .fuzzgen import corpus


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n_tries = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    n_mut = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
    n_weird = int(sys.argv[4]) if len(sys.argv) > 4 else 4000
    n_nested = int(sys.argv[5]) if len(sys.argv) > 5 else 2000
    n_reseal = int(sys.argv[6]) if len(sys.argv) > 6 else 4000
    n_padded = int(sys.argv[7]) if len(sys.argv) > 7 else 1000
    o = Oracle()
    ref = RefElf()
    cases = corpus(seed, o.keccak256, n_tries, n_mut, n_weird, n_nested, n_reseal, n_padded)
    t0 = time.time()

    def one(c):
        a = o.verify(c["root"], c["proof"], c["key"])
        a2 = o.verify(c["root"], c["proof"], c["key"], mirror=True)
        b = ref.run(c["root"], c["proof"], c["key"])
        return a, a2, b

    with ThreadPoolExecutor(8) as ex:
        res = list(ex.map(one, cases))
    bad = 0
    hist = Counter()
    for c, (a, a2, b) in zip(cases, res):
        hist[(c["tag"].split("+")[0], STATUS_NAMES[b["status"]])] += 1
        ok = (a[0] == b["status"]) and (a[1] == b["value"]) and (a2[0] == a[0]) and (a2[1] == a[1])
        if not ok:
            bad += 1
            if bad <= 15:
                print("MISMATCH", c["tag"], "oracle:", a[0], a[1].hex() if a[1] is not None else None,
                      "ref:", b["status"], b["value"].hex() if b["value"] is not None else None)
                print("   root", c["root"].hex(), "key", c["key"].hex())
                for nd in c["proof"]:
                    print("   node", nd.hex())
                print("   stderr:", b["stderr"].split("\n")[1] if "\n" in b["stderr"] else b["stderr"])
    print(f"{len(cases)} cases, {bad} mismatches, {time.time()-t0:.1f}s")
    for k in sorted(hist):
        print("  ", k, hist[k])
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
