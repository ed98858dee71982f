"""Live differential check of the C restatement against the reference's own guest ELF.  Needs
/root/reference (this container only; skipped on the GPU box)."""
import pytest

from oracle.pyoracle import ref_available


@pytest.mark.skipif(not ref_available(), reason="/root/reference not mounted")
def test_fuzz_small(oracle):
    from oracle.fuzzgen import corpus
    from oracle.pyoracle import RefElf
    ref = RefElf()
    cases = corpus(12345, oracle.keccak256, 20, 300, 700, 100, 100, 60)
    bad = []
    for c in cases:
        a = oracle.verify(c["root"], c["proof"], c["key"])
        r = ref.run(c["root"], c["proof"], c["key"])
        if a[0] != r["status"] or a[1] != r["value"]:
            bad.append(c["tag"])
    assert not bad, bad[:10]


@pytest.mark.skipif(not ref_available(), reason="/root/reference not mounted")
def test_appendix_e_step_counts(golden):
    """The emulator reproduces the instruction counts recorded in SURVEY.md Appendix E."""
    from oracle.pyoracle import RefElf
    ref = RefElf()
    want = [85738, 87571, 65858, 84258, 66526, 124355, 124348, 126022, 126099, 94002, 198144, 122067, 159926, 58404]
    vs = [v for v in golden["vectors"] if v["tag"].startswith("appendixE/")]
    assert len(vs) == 14
    for v, w in zip(vs, want):
        r = ref.run(v["root_b"], v["proof_b"], v["key_b"])
        assert r["steps"] == w == v["steps"]
        assert r["status"] == v["status"]
