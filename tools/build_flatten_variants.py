"""libmptv.so variants with different software-prefetch distances of the borsh flattener (host_flatten.h MPTV_STREAM_AHEAD)
into build/variants/:   python tools/build_flatten_variants.py 1024 4096 8192"""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("mptv_build", os.path.join(ROOT, "zk-state-proofs_b200", "build.py"))
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
os.makedirs(os.path.join(ROOT, "build", "variants"), exist_ok=True)
for v in sys.argv[1:]:
    out = os.path.join(ROOT, "build", "variants", f"libmptv_ahead{v}.so")
    print(m.build(extra_flags=(f"-DMPTV_STREAM_AHEAD={v}",), out=out))
