"""CPU restatement of the prover's host glue and three-stage flow -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product
(ligero-prover_b200/, include/) never does.

What is restated (file:line in /root/reference):
  transcript        include/zkp/hash.hpp:47-129,341-346; src/webgpu_prover.cpp:281-282,337-341
                    (a string literal is hashed WITH its NUL terminator: array overload, hash.hpp:61-65)
  hash PRG          include/zkp/random.hpp:87-146 (first block SHA-256(LE64(0)) without the seed, later blocks
                    SHA-256(seed || LE64(i)); bytes consumed from index 31 downward)
  sampler           include/util/portable_sample.hpp:17-33 over boost::random::uniform_int_distribution
                    -- PARITY UNPINNED: Boost is not in the image; restated from the published algorithm
  AES stream        include/util/csprng.hpp:28-110 + include/zkp/finite_field_gmp.hpp:70-78
  rows / masks      include/zkp/backend/witness_manager.hpp:200-336
  stages            include/zkp/nonbatch_context.hpp:445-494,555-558 (1), :654-780 (2), :935-993 (3)
  openings          include/zkp/merkle_tree.hpp:155-318; include/zkp/proof_serializer.hpp:82-117
  container         proto/ligero_proof.proto:13-60, proto/common.proto:21-33 (parsed here with google.protobuf
                    descriptors built at run time: an implementation independent of the product's wire writer)
Field arithmetic, encoding and column hashing come from the C oracle (oracle/lgo.py).
"""
import gzip
import hashlib
import struct

import numpy as np

from . import lgo

P = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


# ---------------------------------------------------------------- transcript / randomness
def stage1_seed(root: bytes, instance_hash: bytes) -> bytes:
    return hashlib.sha256(b"LigetronStage1\x00" + root + instance_hash).digest()


def stage2_seed(root: bytes, code, linear, quad) -> bytes:
    h = hashlib.sha256(b"LigetronStage2\x00" + root)
    for v in (code, linear, quad):
        h.update(np.ascontiguousarray(v, dtype="<u4").tobytes())
    return h.digest()


class HashRandomEngine:
    def __init__(self, seed: bytes):
        self.seed, self.state, self.offset, self.buf, self.prefix = seed, 0, -1, b"", b""

    def __call__(self) -> int:
        if self.offset < 0 or self.offset >= 32:
            self.buf = hashlib.sha256(self.prefix + struct.pack("<Q", self.state)).digest()
            self.state += 1
            self.prefix = self.seed
            self.offset = 31
        b = self.buf[self.offset]
        self.offset -= 1
        return b


def boost_uniform_int(eng, lo: int, hi: int) -> int:
    """boost::random::detail::generate_uniform_int for an 8-bit engine, 64-bit result"""
    rng, brange, M = hi - lo, 255, (1 << 64) - 1
    if rng == 0:
        return lo
    if rng == brange:
        return eng() + lo
    if brange < rng:
        while True:
            if rng == M:
                limit = rng // (brange + 1) + (1 if rng % (brange + 1) == brange else 0)
            else:
                limit = (rng + 1) // (brange + 1)
            result, mult = 0, 1
            while mult <= limit:
                result = (result + eng() * mult) & M
                if (mult * brange) & M == (rng - mult + 1) & M:
                    return result
                mult = (mult * (brange + 1)) & M
            inc = boost_uniform_int(eng, 0, rng // mult)
            if M // mult < inc:
                continue
            inc *= mult
            result = (result + inc) & M
            if result < inc or result > rng:
                continue
            return result + lo
    bucket = (brange // (rng + 1)) & 0xFF
    if brange % (rng + 1) == rng:
        bucket = (bucket + 1) & 0xFF
    while True:
        r = eng() // bucket
        if r <= rng:
            return r + lo


def sample_indices(seed: bytes, n: int, sample_size: int = 192):
    eng = HashRandomEngine(seed)
    idx, out = list(range(n)), []
    for i in range(min(sample_size, n)):
        j = boost_uniform_int(eng, i, n - 1)
        idx[i], idx[j] = idx[j], idx[i]
        out.append(idx[i])
    return sorted(out)


class FrRandomStream:
    """AES-256-CTR keystream -> field elements (32 bytes as LE integer, >> 2, one conditional subtract)"""

    def __init__(self, key: bytes, iv: bytes = bytes(16)):
        from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
        self.enc = Cipher(algorithms.AES(key), modes.CTR(iv)).encryptor()

    def next(self) -> int:
        v = int.from_bytes(self.enc.update(bytes(32)), "little") >> 2
        return v - P if v >= P else v

    def take(self, count: int):
        return [self.next() for _ in range(count)]


# ---------------------------------------------------------------- Merkle openings
def sibling_positions(known, total_count):
    pos, known = [], set(known)
    start, end = total_count // 2, total_count
    while start > 0:
        upper = set()
        for i in range(start, end, 2):
            ll, lr = i - start, i - start + 1
            kl, kr = ll in known, lr in known
            if kl and kr:
                upper.add(ll // 2)
            elif kr:
                pos.append(i); upper.add(ll // 2)
            elif kl:
                pos.append(i + 1); upper.add(ll // 2)
        known = upper
        start, end = (start - 1) // 2, (end - 1) // 2
    return pos


def recommit(leaves, known, total_count, siblings):
    """leaves: {leaf index: 32-byte digest}; siblings in canonical order; plain binary-tree recomputation
    (deliberately NOT the level-set walk of merkle_tree::recommit_helper, to cross-check it)"""
    saved = dict(zip(sibling_positions(known, total_count), siblings))
    half = total_count // 2
    have = {half + i: d for i, d in leaves.items()}

    def node(i):
        if i in have:
            return have[i]
        if i in saved:
            return saved[i]
        if i >= half:
            raise KeyError("leaf %d neither opened nor supplied" % (i - half))
        return hashlib.sha256(node(2 * i + 1) + node(2 * i + 2)).digest()
    return node(0)


# ---------------------------------------------------------------- container
def _proto_classes():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    pool = descriptor_pool.DescriptorPool()

    def field(msg, name, num, typ, label=F.LABEL_OPTIONAL, type_name=None):
        f = msg.field.add(); f.name, f.number, f.type, f.label = name, num, typ, label
        if type_name:
            f.type_name = type_name

    ts = descriptor_pb2.FileDescriptorProto(name="ts.proto", package="google.protobuf", syntax="proto3")
    m = ts.message_type.add(); m.name = "Timestamp"
    field(m, "seconds", 1, F.TYPE_INT64); field(m, "nanos", 2, F.TYPE_INT32)
    pool.Add(ts)
    common = descriptor_pb2.FileDescriptorProto(name="common.proto", package="ligero.common.v1", syntax="proto3")
    m = common.message_type.add(); m.name = "HashDigest"; field(m, "value", 1, F.TYPE_BYTES)
    m = common.message_type.add(); m.name = "MerkleDecommitment"
    field(m, "algorithm", 1, F.TYPE_UINT32)          # enums travel as varints
    field(m, "root", 2, F.TYPE_MESSAGE, type_name=".ligero.common.v1.HashDigest")
    field(m, "sibling_hashes", 3, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".ligero.common.v1.HashDigest")
    field(m, "leaf_indices", 4, F.TYPE_UINT32, F.LABEL_REPEATED)
    pool.Add(common)
    lp = descriptor_pb2.FileDescriptorProto(name="ligero_proof.proto", package="ligero.v1", syntax="proto3",
                                            dependency=["common.proto", "ts.proto"])
    m = lp.message_type.add(); m.name = "FixedU32Vector"; field(m, "values", 1, F.TYPE_FIXED32, F.LABEL_REPEATED)
    m = lp.message_type.add(); m.name = "ProofMetadata"
    field(m, "prover_version", 1, F.TYPE_STRING); field(m, "proof_schema_version", 2, F.TYPE_UINT32)
    field(m, "proof_type", 3, F.TYPE_UINT32)
    field(m, "program_hash", 4, F.TYPE_MESSAGE, type_name=".ligero.common.v1.HashDigest")
    field(m, "generated_at", 5, F.TYPE_MESSAGE, type_name=".google.protobuf.Timestamp")
    for i, nm in enumerate(("packing_size", "codeword_size", "sample_size", "security_level")):
        field(m, nm, 6 + i, F.TYPE_UINT32)
    m = lp.message_type.add(); m.name = "LigeroProof"
    field(m, "merkle_tree", 1, F.TYPE_MESSAGE, type_name=".ligero.common.v1.MerkleDecommitment")
    for i, nm in enumerate(("encoded_code", "encoded_linear", "encoded_quadratic", "sampled_data")):
        field(m, nm, 2 + i, F.TYPE_MESSAGE, type_name=".ligero.v1.FixedU32Vector")
    m = lp.message_type.add(); m.name = "LigeroProofEnvelope"
    field(m, "metadata", 1, F.TYPE_MESSAGE, type_name=".ligero.v1.ProofMetadata")
    field(m, "ligero_proof", 2, F.TYPE_MESSAGE, type_name=".ligero.v1.LigeroProof")
    pool.Add(lp)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("ligero.v1.LigeroProofEnvelope"))


_ENVELOPE = None


def parse_envelope(data: bytes):
    """gzip or plain LigeroProofEnvelope -> google.protobuf message"""
    global _ENVELOPE
    if _ENVELOPE is None:
        _ENVELOPE = _proto_classes()
    if data[:2] == b"\x1f\x8b":
        data = gzip.decompress(data)
    env = _ENVELOPE()
    env.ParseFromString(data)
    return env


def build_envelope(meta: dict, root: bytes, siblings, leaf_indices, code, linear, quad, samplings) -> bytes:
    """the envelope as google.protobuf serialises it (field-number order; deterministic for these types)"""
    global _ENVELOPE
    if _ENVELOPE is None:
        _ENVELOPE = _proto_classes()
    env = _ENVELOPE()
    md = env.metadata
    md.prover_version = meta["prover_version"]; md.proof_schema_version = 1; md.proof_type = 1
    md.program_hash.value = meta["program_hash"]; md.generated_at.seconds = meta["generated_at"]
    md.packing_size, md.codeword_size, md.sample_size, md.security_level = meta["k"], meta["n"], meta["sample_size"], 128
    mt = env.ligero_proof.merkle_tree
    mt.algorithm = 1; mt.root.value = root
    for s in siblings:
        mt.sibling_hashes.add().value = s
    mt.leaf_indices.extend(int(i) for i in leaf_indices)
    for name, v in (("encoded_code", code), ("encoded_linear", linear), ("encoded_quadratic", quad), ("sampled_data", samplings)):
        getattr(env.ligero_proof, name).values.extend(np.ascontiguousarray(v, dtype=np.uint32).reshape(-1).tolist())
    return env.SerializeToString(deterministic=True)


# ---------------------------------------------------------------- the prover, on the CPU
def int_rows(vals):
    return lgo.to_limbs(vals)


def pad_rows(kinds, values, l, k, enc: FrRandomStream):
    """values: [rows, l, 8] in emission order -> [rows, k, 8] with k-l pads per row from the encoding stream"""
    rows = values.shape[0]
    out = np.zeros((rows, k, 8), np.uint32)
    out[:, :l] = values
    for r in range(rows):
        out[r, l:] = lgo.to_limbs(enc.take(k - l))
    return out


def masks(l, k, enc: FrRandomStream):
    """witness_manager.hpp:271-321"""
    mc = np.zeros((k, 8), np.uint32)
    mc[:l] = lgo.to_limbs(enc.take(l))
    ml = [0] * (2 * k)
    for i in range(l - 1):
        ml[2 * i + 1] = enc.next()
    if l >= 1:
        ml[2 * l - 1] = (-sum(ml[1:2 * (l - 1):2])) % P
    for i in range(2 * (k - l)):
        ml[2 * l + i] = enc.next()
    mq = [0] * (2 * k)
    for i in range(l):
        mq[2 * i + 1] = enc.next()
    for i in range(2 * (k - l)):
        mq[2 * l + i] = enc.next()
    return mc, lgo.to_limbs(ml), lgo.to_limbs(mq)


# event kinds (include/lgr_prover.h LGRP_EV_*): 0 / 1 from the scalar backend, 2.. vbn254fr host calls
# (include/host_modules/vbn254fr.hpp:139-566) with their on_batch_* callbacks (nonbatch_context.hpp:497-553,782-847,996-1047)
EV_LINEAR, EV_QUAD, EV_VSET, EV_VCOPY, EV_VADD, EV_VSUB, EV_VMUL, EV_VDIV, EV_VASSERT_EQ, EV_VBIT = range(10)
EV_VADDC, EV_VSUBC, EV_VCSUB, EV_VMULC, EV_VMONTMULC = range(10, 15)


def run_events(l, k, kinds, values, coefs, enc, arena_slots=0, batch_args=None, batch_consts=None):
    """Replays the event list the way the interpreter would drive a stage context: returns the committed rows
    (k elements each, pads included), their coefficient rows, and per event the list of committed-row indices.
    vbn254fr variables are k-element device buffers; every arena operation acts on all k elements."""
    values = np.ascontiguousarray(values, np.uint32).reshape(-1, l, 8)
    coefs = np.zeros_like(values) if coefs is None else np.ascontiguousarray(coefs, np.uint32).reshape(-1, l, 8)
    arena = np.zeros((max(arena_slots, 1), k, 8), np.uint32)
    args = None if batch_args is None else np.ascontiguousarray(batch_args, np.uint32).reshape(-1, 3)
    consts = None if batch_consts is None else np.ascontiguousarray(batch_consts, np.uint32).reshape(-1, 8)
    M, C, ev_rows = [], [], []
    hr = nb = nc = 0

    def commit(row, coef=None):
        M.append(np.array(row, np.uint32)); C.append(np.zeros((k, 8), np.uint32) if coef is None else coef)
        return len(M) - 1

    for kind in kinds:
        kind = int(kind)
        if kind in (EV_LINEAR, EV_QUAD):
            idx = []
            for _ in range(3 if kind == EV_QUAD else 1):
                row = np.zeros((k, 8), np.uint32); row[:l] = values[hr]
                row[l:] = lgo.to_limbs(enc.take(k - l))                       # witness_manager.hpp:200-214
                cf = np.zeros((k, 8), np.uint32); cf[:l] = coefs[hr]
                idx.append(commit(row, cf)); hr += 1
            ev_rows.append(idx)
            continue
        out, x, y = (int(v) for v in args[nb]); nb += 1
        if kind == EV_VSET:                                                   # write_buffer_clear + on_batch_init
            arena[out] = 0; arena[out, :l] = values[hr]; hr += 1
            arena[out, l:] = lgo.to_limbs(enc.take(k - l))
            ev_rows.append([commit(arena[out])])
        elif kind == EV_VCOPY:                                                # on_batch_equal(out, in)
            arena[out] = arena[x]
            ev_rows.append([commit(arena[out]), commit(arena[x])])
        elif kind == EV_VASSERT_EQ:                                           # on_batch_equal(x, y): operands are args 0, 1
            ev_rows.append([commit(arena[out]), commit(arena[x])])
        elif kind == EV_VADD:
            arena[out] = lgo.elt_add(arena[x], arena[y]); ev_rows.append([])
        elif kind == EV_VSUB:
            arena[out] = lgo.elt_sub(arena[x], arena[y]); ev_rows.append([])
        elif kind == EV_VMUL:                                                 # on_batch_quadratic(x, y, tmp)
            tmp = lgo.elt_mul(arena[x], arena[y])
            ev_rows.append([commit(arena[x]), commit(arena[y]), commit(tmp)])
            arena[out] = tmp
        elif kind == EV_VDIV:                                                 # on_batch_quadratic(tmp, y, x)
            tmp = lgo.elt_div(arena[x], arena[y])
            ev_rows.append([commit(tmp), commit(arena[y]), commit(arena[x])])
            arena[out] = tmp
        elif kind == EV_VBIT:                                                 # on_batch_bit(out)
            arena[out] = lgo.elt_bit(arena[x], y)
            ev_rows.append([commit(arena[out])])
        else:
            c = lgo.from_limbs(consts[nc])[0]; nc += 1
            fn = {EV_VADDC: lgo.elt_add_const, EV_VSUBC: lgo.elt_sub_const, EV_VCSUB: lgo.elt_const_sub, EV_VMULC: lgo.elt_mul_const,
                  EV_VMONTMULC: lgo.elt_montmul_const}[kind]
            arena[out] = fn(arena[x], c); ev_rows.append([])
    return M, C, ev_rows


def stage2_combine(kinds, ev_rows, rowvec, coefvec, s1, shape):
    """code / linear / quad accumulators over `rowvec[r]` (encoded rows, or their sampled columns) in callback order"""
    code_rng, quad_rng = FrRandomStream(s1), FrRandomStream(s1)
    code = np.zeros(shape, np.uint32); linear = np.zeros(shape, np.uint32); quad = np.zeros(shape, np.uint32)
    for kind, idx in zip(kinds, ev_rows):
        kind = int(kind)
        if kind in (EV_LINEAR, EV_QUAD):
            for r in idx:
                code = lgo.elt_fma_const(code, rowvec[r], code_rng.next())
            for r in idx:
                linear = lgo.elt_fma(linear, rowvec[r], coefvec(r))
        elif kind in (EV_VSET, EV_VBIT, EV_VMUL, EV_VDIV):                    # check_code on every committed row of the event
            for r in idx:
                code = lgo.elt_fma_const(code, rowvec[r], code_rng.next())
        if kind in (EV_QUAD, EV_VMUL, EV_VDIV):
            t = lgo.elt_sub(lgo.elt_mul(rowvec[idx[0]], rowvec[idx[1]]), rowvec[idx[2]])
            quad = lgo.elt_fma_const(quad, t, quad_rng.next())
        elif kind == EV_VBIT:                                                 # y := x, z := x, then check_quadratic
            x = rowvec[idx[0]]
            quad = lgo.elt_fma_const(quad, lgo.elt_sub(lgo.elt_mul(x, x), x), quad_rng.next())
        elif kind in (EV_VCOPY, EV_VASSERT_EQ):                               # EltwiseSubMod then EltwiseFMAMod(r)
            quad = lgo.elt_fma_const(quad, lgo.elt_sub(rowvec[idx[0]], rowvec[idx[1]]), quad_rng.next())
    return code, linear, quad


def prove(l, k, kinds, values, coefs, const_sum, encoding_seed, instance_hash, sample_size=192, arena_slots=0, batch_args=None,
          batch_consts=None):
    """Returns a dict with every value that goes into the proof plus the self-check flags."""
    n = 4 * k
    enc = FrRandomStream(encoding_seed)
    M, C, ev_rows = run_events(l, k, kinds, values, coefs, enc, arena_slots, batch_args, batch_consts)
    rows = len(M)
    mc, ml, mq = masks(l, k, enc)
    cw = [lgo.encode(M[r], k) for r in range(rows)]
    cw_masks = [lgo.encode(mc, k), lgo.encode_2k(ml, k), lgo.encode_2k(mq, k)]
    # stage 1
    sha = lgo.Sha(n); sha.init()
    for e in cw + cw_masks:
        sha.update(e)
    digests = sha.final()
    nodes = lgo.merkle_build(digests)
    root = nodes[0].tobytes()
    s1 = stage1_seed(root, instance_hash)
    # stage 2
    code, linear, quad = stage2_combine(kinds, ev_rows, cw, lambda r: lgo.encode(C[r], k), s1, (n, 8))
    code = lgo.elt_add_assign(code, cw_masks[0]); linear = lgo.elt_add_assign(linear, cw_masks[1]); quad = lgo.elt_add_assign(quad, cw_masks[2])
    s2 = stage2_seed(root, code, linear, quad)
    sample = sample_indices(s2, n, sample_size)
    total = nodes.shape[0]
    pos = sibling_positions(sample, total)
    siblings = [nodes[p].tobytes() for p in pos]
    # self-check
    dc, dl, dq = lgo.decode(code, k), lgo.decode(linear, k), lgo.decode(quad, k)
    valid_code = not dc[k:].any()
    valid_linear = (sum(lgo.from_limbs(dl[:l])) + const_sum) % P == 0
    valid_quad = not dq[:l].any()
    # stage 3
    samplings = np.stack([e[sample] for e in cw + cw_masks])
    return {"root": root, "stage1_seed": s1, "stage2_seed": s2, "code": code, "linear": linear, "quad": quad, "sample": sample,
            "positions": pos, "siblings": siblings, "samplings": samplings, "digests": digests, "total_count": total,
            "valid": (valid_code, valid_linear, valid_quad), "encoded_rows": rows + 3}


def verify_openings(proof_env, l, k, kinds, coefs, instance_hash):
    """(vbn254fr events need no operands here: the verifier sees only committed rows, in event order.)
    The checks of src/webgpu_verifier.cpp:314-315,412-442 that need no re-execution: the sampled columns
    re-hash to leaves that recommit to the root, and the three test vectors at the sampled positions equal
    the combinations recomputed from the opened columns (with r re-derived from the transcript)."""
    n = 4 * k
    pr = proof_env.ligero_proof
    root = pr.merkle_tree.root.value
    sample = list(pr.merkle_tree.leaf_indices)
    S = len(sample)
    code = np.array(pr.encoded_code.values, np.uint32).reshape(n, 8)
    linear = np.array(pr.encoded_linear.values, np.uint32).reshape(n, 8)
    quad = np.array(pr.encoded_quadratic.values, np.uint32).reshape(n, 8)
    samp = np.array(pr.sampled_data.values, np.uint32).reshape(-1, S, 8)
    rows = samp.shape[0] - 3
    assert sample == sample_indices(stage2_seed(root, code, linear, quad), n, proof_env.metadata.sample_size), "sample indices do not follow from the transcript"
    # leaves of the opened columns (the verifier hashes the sampled rows with S instances, nonbatch_context.hpp:1097-1104)
    sha = lgo.Sha(S); sha.init()
    for r in range(rows + 3):
        sha.update(samp[r])
    leaf_digests = sha.final()
    leaves = {sample[i]: leaf_digests[i].tobytes() for i in range(S)}
    total = 2 * (1 << (n - 1).bit_length()) - 1
    assert recommit(leaves, sample, total, [s.value for s in pr.merkle_tree.sibling_hashes]) == root, "openings do not recommit to the root"
    s1 = stage1_seed(root, instance_hash)
    # which committed rows belong to which event, and the coefficient rows of the scalar events
    committed = {EV_LINEAR: 1, EV_QUAD: 3, EV_VSET: 1, EV_VBIT: 1, EV_VCOPY: 2, EV_VASSERT_EQ: 2, EV_VMUL: 3, EV_VDIV: 3}
    coefs = np.ascontiguousarray(coefs, np.uint32).reshape(-1, l, 8)
    ev_rows, coef_of, r, hr = [], {}, 0, 0
    for kind in kinds:
        kind = int(kind)
        cnt = committed.get(kind, 0)
        ev_rows.append(list(range(r, r + cnt)))
        if kind in (EV_LINEAR, EV_QUAD):
            for j in range(cnt):
                coef_of[r + j] = hr; hr += 1
        elif kind == EV_VSET:
            hr += 1
        r += cnt
    assert r == rows, "the proof holds %d rows, the statement commits %d" % (rows, r)

    def coefvec(row):
        crow = np.zeros((k, 8), np.uint32); crow[:l] = coefs[coef_of[row]]
        return lgo.encode(crow, k)[sample]
    acc_c, acc_l, acc_q = stage2_combine(kinds, ev_rows, samp, coefvec, s1, (S, 8))
    acc_c = lgo.elt_add_assign(acc_c, samp[rows]); acc_l = lgo.elt_add_assign(acc_l, samp[rows + 1]); acc_q = lgo.elt_add_assign(acc_q, samp[rows + 2])
    assert np.array_equal(acc_c, code[sample]) and np.array_equal(acc_l, linear[sample]) and np.array_equal(acc_q, quad[sample]), "test vectors disagree with the openings"
    return True


# ---------------------------------------------------------------- row packing (witness_manager.hpp:117-269,497-503)
def pack_rows(l, witnesses):
    """witnesses: sequence of ("L", value, coef) / ("Q", (x, y, z), (cx, cy, cz)) in release order (python ints).
    Returns (kinds, values, coefs) with rows as lists of l ints -- lazy emission: a row leaves when a witness arrives
    and the open row is full; finalize emits the partial linear row, then the partial triple."""
    kinds, values, coefs = [], [], []
    lin_v, lin_c = [], []
    qv, qc = ([], [], []), ([], [], [])

    def pad(r):
        return r + [0] * (l - len(r))

    def flush_lin():
        if lin_v:
            kinds.append(0); values.append(pad(list(lin_v))); coefs.append(pad(list(lin_c)))
            lin_v.clear(); lin_c.clear()

    def flush_quad():
        if qv[0]:
            kinds.append(1)
            for i in range(3):
                values.append(pad(list(qv[i]))); coefs.append(pad(list(qc[i])))
                qv[i].clear(); qc[i].clear()

    for w in witnesses:
        if w[0] == "L":
            if len(lin_v) >= l:
                flush_lin()
            lin_v.append(w[1]); lin_c.append(w[2])
        else:
            if len(qv[0]) >= l:
                flush_quad()
            for i in range(3):
                qv[i].append(w[1][i]); qc[i].append(w[2][i])
    flush_lin(); flush_quad()
    return kinds, values, coefs
