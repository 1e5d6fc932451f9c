"""Multi-GPU commitment layout (DESIGN.md "Multi-GPU").

Rows shard contiguously over ranks: rank g owns rows [g*R, (g+1)*R) and commits them exactly as a
single GPU would (encode every row, one SHA-256 stream per codeword column over ITS rows).  The only
exchange is one all-gather of the n leaf digests per rank; every rank then builds the same Merkle
tree over the G*n leaves, leaf index = g*n + j.  With G = 1 this is the reference's commitment
(include/zkp/nonbatch_context.hpp:555-558 + include/zkp/merkle_tree.hpp:343-375) bit for bit; with
G > 1 it is the sharded commitment BASELINE.json's north_star describes ("a single NCCL all-gather
... only to assemble the Merkle root").  A column opening then carries G leaves instead of one.

The functions here are the host-side logic shared by bench.py and the tests; the collective is
whatever backend torch.distributed was initialised with (nccl on GPUs, gloo in the CPU tests).
"""


def shard_rows(total_rows, world, rank):
    """contiguous row range [begin, end) of `rank`; the first total_rows % world ranks get one more"""
    base, extra = divmod(total_rows, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def leaf_index(rank, column, n):
    return rank * n + column


def gather_leaf_digests(local_digests, world, dist=None):
    """all-gather of the per-rank [n, 8] int32 digest tensors -> [world*n, 8] (rank-major)"""
    import torch
    if world == 1:
        return local_digests
    out = torch.empty((world * local_digests.shape[0],) + tuple(local_digests.shape[1:]), dtype=local_digests.dtype, device=local_digests.device)
    dist.all_gather_into_tensor(out, local_digests.contiguous())
    return out
