"""Multi-GPU commitment layouts (DESIGN.md "Multi-GPU").  One process per GPU, torch.distributed for
the plumbing (nccl on GPUs; the CPU tests run the same orchestration on gloo with an oracle engine).

Encoding shards by rows, but a column's leaf is ONE SHA-256 stream over all rows in order
(include/zkp/nonbatch_context.hpp:445-451 -> shader/sha256.wgsl:147-177), so two layouts exist:

1. `commit_sharded` -- what BASELINE.json's north_star describes: rank g commits its own contiguous
   row shard exactly as a single GPU would; the only exchange is ONE all-gather of the n leaf digests
   per rank, and every rank builds the tree over the G*n leaves (leaf index g*n + j).  G = 1 is the
   reference's commitment bit for bit; G > 1 commits to the same data with G leaves per column.

2. `commit_exact` -- bit-exact with the single-GPU / reference root for the WHOLE matrix:
   tiles of T rows are dealt round-robin (global tile t lives on rank t mod G); every round each rank
   encodes its tile, an all-to-all moves the column slab [h*n/G, (h+1)*n/G) of every tile to rank h,
   and rank h absorbs the G received slabs in rank order (= global row order) into its n/G column
   hashes.  A final all-gather of the n/G digests per rank assembles the n leaves; every rank builds
   the reference's tree.  Encoding of round r+1 overlaps the exchange + hashing of round r.
"""


def shard_rows(total_rows, world, rank):
    """contiguous row range [begin, end) of `rank`; the first total_rows % world ranks get one more"""
    base, extra = divmod(total_rows, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def leaf_index(rank, column, n):
    return rank * n + column


def gather_leaf_digests(local_digests, world, dist=None):
    """all-gather of the per-rank [m, 8] int32 digest tensors -> [world*m, 8] (rank-major)"""
    import torch
    if world == 1:
        return local_digests
    out = torch.empty((world * local_digests.shape[0],) + tuple(local_digests.shape[1:]), dtype=local_digests.dtype, device=local_digests.device)
    dist.all_gather_into_tensor(out, local_digests.contiguous())
    return out


def tile_owner(tile, world):
    return tile % world


def tiles_of_rank(num_tiles, world, rank):
    return list(range(rank, num_tiles, world))


class GpuEngine:
    """the device side of commit_exact on one GPU, through the executor (C ABI underneath); slabs travel by ONE
    NCCL all-to-all per round.  The encoder writes the tile slab-major (lgr_encode_rows_slabs), so there is no pack pass.
    Kept as the transport of last resort (CUDA IPC unavailable) and as the reference point for PeerStoreEngine."""
    transport = "nccl-all-to-all"

    def __init__(self, ex, tile_rows, world, rank=0):
        import torch
        self.torch = torch
        self.ex = ex
        self.T = tile_rows
        self.G = world
        self.rank = rank
        self.n = ex.encoding_size()
        self.k = ex.padding_size()
        assert self.n % world == 0 and (self.n // world) % 32 == 0, "n/G must be a multiple of 32 columns"
        self.slab = self.n // world
        dev = "cuda:%d" % ex._device
        words = tile_rows * self.n * 8
        self.send = [torch.empty(words, dtype=torch.int32, device=dev) for _ in range(2)]       # [G][T][n/G]
        self.recv = [torch.empty(words, dtype=torch.int32, device=dev) for _ in range(2)] if world > 1 else self.send
        self.enc_stream = torch.cuda.Stream(device=dev)
        self.hash_stream = torch.cuda.Stream(device=dev, priority=int(__import__("os").environ.get("LGR_EXACT_HASH_PRIORITY", "-1")))   # high: with the gate, the hash CTAs take their SMs first
        self.enc_done = [torch.cuda.Event() for _ in range(2)]
        self.hash_done = [torch.cuda.Event() for _ in range(2)]
        ex.sha256_init(self.slab)
        self.sha_ctx = ex.make_device_buffer(ex.sha256_context_bytes(self.slab))
        self.sha_dig = ex.make_device_buffer(self.slab * 32)
        self.bind = ex.bind_sha256_context(self.sha_ctx, self.sha_dig)

    def begin(self):
        t = self.torch
        self.fork = t.cuda.Event()
        self.fork.record(t.cuda.current_stream())
        self.enc_stream.wait_event(self.fork)
        self.hash_stream.wait_event(self.fork)
        with t.cuda.stream(self.hash_stream):
            self.ex.use_torch_stream()
            self.ex.sha256_init(self.slab)
            self.ex.sha256_digest_init(self.bind)

    def encode_round(self, rnd, rows_buf, nrows):
        """encode this rank's tile of round `rnd` (nrows <= T valid rows) straight into the slab-major send buffer:
        chunk h = column slab h of every row, row-major [T][n/G]"""
        t = self.torch
        b = rnd & 1
        with t.cuda.stream(self.enc_stream):
            if rnd >= 2:
                self.enc_stream.wait_event(self.hash_done[b])          # send[b] free again
            self.ex.use_torch_stream()
            if nrows:
                base = self.send[b].data_ptr()
                self.ex.encode_rows_slabs(rows_buf, nrows, [base + h * self.T * self.slab * 32 for h in range(self.G)])
            self.enc_done[b].record(self.enc_stream)

    def exchange_and_hash(self, rnd, rows_per_rank, dist):
        """all-to-all of the slabs, then absorb the G tiles of this round in global row order"""
        t = self.torch
        b = rnd & 1
        with t.cuda.stream(self.hash_stream):
            self.hash_stream.wait_event(self.enc_done[b])
            if self.G > 1:
                dist.all_to_all_single(self.recv[b], self.send[b])
                src = self.recv[b]
            else:
                src = self.send[b]
            self.ex.use_torch_stream()
            self.ex.sha256_init(self.slab)
            if all(r == self.T for r in rows_per_rank):
                # full round: the G chunks are one contiguous [G*T][n/G] matrix in global row order -> ONE launch.  (With
                # one launch per chunk the encoder of the next round refills the SMs in the gap between two launches and
                # the SM-owning chain CTAs of the next chunk wait for it to drain: tools/exact_round_sim.py.)
                self.ex.sha256_digest_update_rows(self.bind, self.ex.wrap(src), self.G * self.T, self.slab)
            else:
                rv = src.view(self.G, self.T * self.slab * 8)
                for h in range(self.G):
                    if rows_per_rank[h]:
                        self.ex.sha256_digest_update_rows(self.bind, self.ex.wrap(rv[h]), rows_per_rank[h], self.slab)
            self.hash_done[b].record(self.hash_stream)

    def finish(self, dist):
        """final + all-gather of the n/G digests per rank -> [n, 8] int32 leaf digests on every rank"""
        t = self.torch
        with t.cuda.stream(self.hash_stream):
            self.ex.use_torch_stream()
            self.ex.sha256_init(self.slab)
            self.ex.sha256_digest_final(self.bind)
            local = self.sha_dig.storage[: self.slab * 8].view(self.slab, 8)
            leaves = gather_leaf_digests(local, self.G, dist)
            done = t.cuda.Event()
            done.record(self.hash_stream)
        t.cuda.current_stream().wait_event(done)
        self.ex.use_torch_stream()
        return leaves

    def close(self):
        pass


class PeerStoreEngine(GpuEngine):
    """commit_exact with NO data-path collective: the encoder of rank g stores column slab h of its tile straight into
    rank h's receive buffer over NVLink (peer memory through CUDA IPC; the stores are the encoder's own 256-bit codeword
    stores, csrc/kernels.h CodewordSink), and two tiny flag kernels hand the buffers over (csrc/peer_kernels.cu):

        encode stream:  wait consumed[*] >= q-2  ->  encode_rows_slabs(-> peers' recv[q&1])  ->  ready[g] := q  on every peer
        hash stream  :  wait ready[*]    >= q    ->  absorb chunks 0..G-1 of recv[q&1]       ->  consumed[h] := q on every peer

    q counts rounds over the life of the engine (monotone flags, never reset), so commitments can follow each other
    without a barrier.  Per rank one IPC allocation: recv[2][G][T][n/G] | ready[G] u64 | consumed[G] u64 | err u32."""
    transport = "peer-store (CUDA IPC over NVLink)"

    def __init__(self, ex, tile_rows, world, rank, dist):
        import torch
        self.torch = torch
        self.ex = ex
        self.T, self.G, self.rank = tile_rows, world, rank
        self.n, self.k = ex.encoding_size(), ex.padding_size()
        assert self.n % world == 0 and (self.n // world) % 32 == 0, "n/G must be a multiple of 32 columns"
        assert world <= 8
        self.slab = self.n // world
        dev = "cuda:%d" % ex._device
        self.buf_bytes = tile_rows * self.n * 32                     # one receive buffer: [G][T][n/G]
        self.off_ready = 2 * self.buf_bytes
        self.off_consumed = self.off_ready + 64
        self.off_err = self.off_consumed + 64
        # every rank must end up on the same transport: agree after each step that can fail locally
        self.local_ptr, handle, self.peer = None, None, []
        try:
            self.local_ptr, handle = ex.ipc_alloc(self.off_err + 64)
        except Exception as e:                                        # noqa: BLE001 -- reported through the agreement below
            self.fail = "ipc_alloc: %s" % e
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        ok = all(h is not None for h in handles)
        if ok:
            try:
                for h in range(world):
                    self.peer.append(self.local_ptr if h == rank else ex.ipc_open(handles[h]))
            except Exception as e:                                    # noqa: BLE001
                self.fail = "ipc_open: %s" % e
                ok = False
        oks = [None] * world
        dist.all_gather_object(oks, ok)
        if not all(oks):
            self.close()
            raise RuntimeError(getattr(self, "fail", "a peer could not map this rank's memory"))
        self.q = 0                                                    # rounds issued so far (global, monotone)
        self.enc_stream = torch.cuda.Stream(device=dev)
        self.hash_stream = torch.cuda.Stream(device=dev, priority=int(__import__("os").environ.get("LGR_EXACT_HASH_PRIORITY", "-1")))   # high: with the gate, the hash CTAs take their SMs first
        ex.sha256_init(self.slab)
        self.sha_ctx = ex.make_device_buffer(ex.sha256_context_bytes(self.slab))
        self.sha_dig = ex.make_device_buffer(self.slab * 32)
        self.bind = ex.bind_sha256_context(self.sha_ctx, self.sha_dig)
        self.err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        # encode(r+1) is released together with hash(r) (when every rank's slabs of round r have arrived), not earlier:
        # otherwise its grid has filled every SM by then and the SM-owning chain CTAs of the hash wait for it to drain
        # (only where the hash is the SM-owning chain kernel, n/G <= 18944 columns: csrc/sha_kernels.cu launch_sha_update)
        self.gate = __import__("os").environ.get("LGR_EXACT_GATE", "1") != "0" and self.slab <= 18944
        self.hash_go = [torch.cuda.Event() for _ in range(2)]

    def begin(self):
        super().begin()
        self.q0 = self.q

    def encode_round(self, rnd, rows_buf, nrows):
        t = self.torch
        q = self.q0 + rnd + 1
        b = q & 1
        with t.cuda.stream(self.enc_stream):
            self.ex.use_torch_stream()
            if q > 2:                                                 # every peer has hashed what round q-2 left in its recv[b]
                self.ex.peer_wait(self.local_ptr + self.off_consumed, self.G, q - 2, self.local_ptr + self.off_err)
            if self.gate and rnd >= 1:
                self.enc_stream.wait_event(self.hash_go[(rnd - 1) & 1])
            if nrows:
                chunk = b * self.buf_bytes + self.rank * self.T * self.slab * 32
                self.ex.encode_rows_slabs(rows_buf, nrows, [self.peer[h] + chunk for h in range(self.G)])
            self.ex.peer_signal([self.peer[h] + self.off_ready + 8 * self.rank for h in range(self.G)], q)

    def exchange_and_hash(self, rnd, rows_per_rank, dist):
        t = self.torch
        q = self.q0 + rnd + 1
        b = q & 1
        self.q = q
        with t.cuda.stream(self.hash_stream):
            self.ex.use_torch_stream()
            self.ex.sha256_init(self.slab)
            self.ex.peer_wait(self.local_ptr + self.off_ready, self.G, q, self.local_ptr + self.off_err)
            self.hash_go[rnd & 1].record(self.hash_stream)
            if all(r == self.T for r in rows_per_rank):             # full round: one [G*T][n/G] matrix, one launch (see GpuEngine)
                self.ex.sha256_digest_update_rows(self.bind, RawSlice(self.local_ptr + b * self.buf_bytes, self.buf_bytes), self.G * self.T, self.slab)
            else:
                for g in range(self.G):
                    if rows_per_rank[g]:
                        chunk = RawSlice(self.local_ptr + b * self.buf_bytes + g * self.T * self.slab * 32, self.T * self.slab * 32)
                        self.ex.sha256_digest_update_rows(self.bind, chunk, rows_per_rank[g], self.slab)
            self.ex.peer_signal([self.peer[g] + self.off_consumed + 8 * self.rank for g in range(self.G)], q)

    def finish(self, dist):
        t = self.torch
        tail = t.cuda.Event()
        tail.record(self.enc_stream)                                   # the last ready-signal of this rank
        self.hash_stream.wait_event(tail)
        leaves = super().finish(dist)
        # a timed-out hand-over leaves a non-zero code in err: surface it instead of returning a wrong root
        self.ex.read_into(self.err_host, self.local_ptr + self.off_err, 4)
        if int(self.err_host[0]) != 0:
            raise RuntimeError("peer hand-over timed out (flag slot %d) on rank %d" % (int(self.err_host[0]) - 1, self.rank))
        return leaves

    def close(self):
        for h, p in enumerate(self.peer):
            if h != self.rank:
                self.ex.ipc_close(p)
        if self.local_ptr:
            self.ex.ipc_free(self.local_ptr)
        self.peer, self.local_ptr = [], None


class RawSlice:
    """a window of raw device memory (peer-visible IPC allocation) with the Buffer interface the executor needs"""

    def __init__(self, addr, nbytes):
        self.addr, self.nbytes = addr, nbytes

    def ptr(self):
        import ctypes as C
        return C.c_void_p(self.addr)

    def size(self):
        return self.nbytes


def make_gpu_engine(ex, tile_rows, world, rank, dist, transport=None):
    """PeerStoreEngine when the ranks can map each other's memory (one NVLink/NVSwitch box), else the NCCL engine.
    LGR_EXACT_TRANSPORT=nccl forces the collective path (A/B measurements)."""
    import os
    want = transport or os.environ.get("LGR_EXACT_TRANSPORT", "peer")
    if world > 1 and want != "nccl":
        try:
            return PeerStoreEngine(ex, tile_rows, world, rank, dist)
        except RuntimeError as e:                                    # IPC refused on some rank (all ranks raise together)
            import sys
            if rank == 0:
                print("sharding: peer-store transport unavailable (%s); using the NCCL all-to-all" % e, file=sys.stderr, flush=True)
    return GpuEngine(ex, tile_rows, world, rank)


def commit_exact(engine, local_tiles, total_rows, tile_rows, world, rank, dist):
    """Drive `engine` through the exact layout.

    local_tiles: callable(local_tile_index) -> (rows handle for the engine, nrows) for this rank's
    tiles in order (global tile = local_index*world + rank).  total_rows: rows of the global matrix.
    Returns the [n, 8] leaf digests (identical on every rank, identical to a single-GPU commit)."""
    num_tiles = (total_rows + tile_rows - 1) // tile_rows
    rounds = (num_tiles + world - 1) // world
    engine.begin()

    def rows_in_tile(tile):
        if tile >= num_tiles:
            return 0
        return min(tile_rows, total_rows - tile * tile_rows)

    for rnd in range(rounds):
        mine = rnd * world + rank
        handle, nrows = local_tiles(rnd) if mine < num_tiles else (None, 0)
        assert nrows == rows_in_tile(mine)
        engine.encode_round(rnd, handle, nrows)
        engine.exchange_and_hash(rnd, [rows_in_tile(rnd * world + h) for h in range(world)], dist)
    return engine.finish(dist)
