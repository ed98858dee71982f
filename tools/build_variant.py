"""a libmptv.so variant with extra compiler flags into build/variants/:  python tools/build_variant.py NAME -DFLAG[=V] ..."""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("mptv_build", os.path.join(ROOT, "zk-state-proofs_b200", "build.py"))
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
os.makedirs(os.path.join(ROOT, "build", "variants"), exist_ok=True)
print(m.build(extra_flags=tuple(sys.argv[2:]), out=os.path.join(ROOT, "build", "variants", f"libmptv_{sys.argv[1]}.so")))
