import gzip, json, sys
sys.path.insert(0,'.')
import zk_state_proofs_b200 as z
vs=json.loads(gzip.open('tests/golden/verify_vectors.json.gz').read())['vectors']
ver=z.Verifier([0])
inputs=[z.MerkleProofInput([bytes.fromhex(n) for n in v['proof']], bytes.fromhex(v['root']), bytes.fromhex(v['key'])) for v in vs]
for fused in (1,0):
  for lanes in (8,32):
    ver.set_option('fused_classify',fused); ver.set_option('lanes_per_proof',lanes)
    res=ver.verify_merkle_proofs(inputs)
    bad=[i for i,(v,r) in enumerate(zip(vs,res)) if (r.status if isinstance(r,z.VerifyPanic) else 0)!=v['status']]
    print('fused',fused,'lanes',lanes,'bad',bad)
for i in bad[:2]:
    v=vs[i]; print(v['tag'], v['status'], v['key'], v['root']); [print('  ',n) for n in v['proof']]
