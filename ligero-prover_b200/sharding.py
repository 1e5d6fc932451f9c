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
    """the device side of commit_exact on one GPU, through the executor (C ABI underneath)"""

    def __init__(self, ex, tile_rows, world):
        import torch
        self.torch = torch
        self.ex = ex
        self.T = tile_rows
        self.G = world
        self.n = ex.encoding_size()
        self.k = ex.padding_size()
        assert self.n % world == 0 and (self.n // world) % 32 == 0, "n/G must be a multiple of 32 columns"
        self.slab = self.n // world
        dev = "cuda:%d" % ex._device
        words = tile_rows * self.n * 8
        self.tile = [torch.empty(words, dtype=torch.int32, device=dev) for _ in range(2)]       # [T][n] codewords
        self.send = [torch.empty(words, dtype=torch.int32, device=dev) for _ in range(2)]       # [G][T][n/G]
        self.recv = [torch.empty(words, dtype=torch.int32, device=dev) for _ in range(2)]       # [G][T][n/G] = [G*T][n/G]
        self.enc_stream = torch.cuda.Stream(device=dev)
        self.hash_stream = torch.cuda.Stream(device=dev)
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
        """encode this rank's tile of round `rnd` (nrows <= T valid rows; the rest of the tile is the
        encoding of zero rows = zero codewords, never hashed) and pack it slab-major"""
        t = self.torch
        b = rnd & 1
        with t.cuda.stream(self.enc_stream):
            if rnd >= 2:
                self.enc_stream.wait_event(self.hash_done[b])          # send[b] / tile[b] free again
            self.ex.use_torch_stream()
            tile = self.ex.wrap(self.tile[b])
            if nrows:
                self.ex.encode_rows(rows_buf, nrows, tile)
            tv = self.tile[b].view(self.T, self.n * 8)
            sv = self.send[b].view(self.G, self.T, self.slab * 8)
            for h in range(self.G):                                     # column slab h of every row -> chunk h
                sv[h, :nrows].copy_(tv[:nrows, h * self.slab * 8:(h + 1) * self.slab * 8], non_blocking=True)
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
