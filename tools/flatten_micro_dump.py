"""dump a config-2 style borsh batch for tools/flatten_micro.cpp:  python tools/flatten_micro_dump.py [n] [accounts]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from workload import gen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
acc = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
trie = gen.SynthTrie(acc, 2, kind=0)
b = gen.account_batch(trie, n, seed=2)
blobs, boff = gen.batch_to_borsh(b)
blobs.tofile("/tmp/blobs.bin"); boff.tofile("/tmp/boff.bin")
print(len(blobs), len(boff))
