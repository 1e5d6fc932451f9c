"""Generates tests/golden/prover_k64.json: a small statement and every value of its proof as computed by the CPU
restatement (oracle/prover_ref.py) -- NOT by running the reference, which cannot be built here.
The fixture pins the restatement against regressions and lets the GPU prover be compared with committed bytes.
Run from the repo root:  python tests/golden/make_prover_fixture.py"""
import hashlib, json, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import lgo, prover_ref as ref

k, l, kinds = 64, 40, [0, 1, 0, 1]
rng = random.Random(20261017)
vals, coefs, acc = [], [], 0
for kind in kinds:
    if kind:
        x = [rng.randrange(ref.P) for _ in range(l)]; y = [rng.randrange(ref.P) for _ in range(l)]
        rows = [x, y, [a * b % ref.P for a, b in zip(x, y)]]
    else:
        rows = [[rng.randrange(1 << 64) for _ in range(l)]]
    for r in rows:
        c = [rng.randrange(ref.P) for _ in range(l)]
        acc += sum(a * b for a, b in zip(r, c)); vals.append(r); coefs.append(c)
values = np.stack([lgo.to_limbs(r) for r in vals]); coef = np.stack([lgo.to_limbs(c) for c in coefs])
const_sum = (-acc) % ref.P
enc_seed = hashlib.sha256(b"golden encoding seed").digest()
inst = hashlib.sha256(b"golden instance").digest()
prog = hashlib.sha256(b"golden program").digest()
w = ref.prove(l, k, kinds, values, coef, const_sum, enc_seed, inst)
meta = {"prover_version": "1.5.0", "program_hash": prog, "generated_at": 1792214281, "k": k, "n": 4 * k, "sample_size": 192}
env = ref.build_envelope(meta, w["root"], w["siblings"], w["sample"], w["code"], w["linear"], w["quad"], w["samplings"])
out = {
    "k": k, "l": l, "kinds": kinds, "values_hex": values.tobytes().hex(), "coefs_hex": coef.tobytes().hex(), "const_sum": hex(const_sum),
    "encoding_seed": enc_seed.hex(), "instance_hash": inst.hex(), "program_hash": prog.hex(), "generated_at": 1792214281,
    "root": w["root"].hex(), "stage1_seed": w["stage1_seed"].hex(), "stage2_seed": w["stage2_seed"].hex(), "sample": w["sample"],
    "envelope_sha256": hashlib.sha256(env).hexdigest(), "envelope_len": len(env), "valid": list(w["valid"]),
    "code_sha256": hashlib.sha256(w["code"].tobytes()).hexdigest(), "samplings_sha256": hashlib.sha256(w["samplings"].tobytes()).hexdigest(),
}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "prover_k64.json"), "w"), indent=1)
print("root", out["root"], "envelope", out["envelope_len"], "bytes", out["envelope_sha256"])
