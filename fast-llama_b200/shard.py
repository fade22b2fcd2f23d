"""Request-batch sharding across GPUs: weights are replicated, sequence i of a batch of B lives on rank
i // (B / world) in KV slot i % (B / world); the only traffic between ranks is one all-gather of the sampled int32 tokens per
decode step, so that every rank (and the host that prints / detects stops) sees every sequence's token.  With B == 1 there is
one rank and no collective.  Pure host logic: works with the gloo backend on CPU tensors (tests) and NCCL on device tensors."""
import torch
import torch.distributed as dist


def local_sequences(batch: int, world: int, rank: int):
    """indices of the sequences this rank owns (contiguous, batch must divide evenly like the reference's max_batch_size slots)"""
    if batch % world:
        raise ValueError(f"batch {batch} is not a multiple of the world size {world}")
    per = batch // world
    return list(range(rank * per, (rank + 1) * per))


def owner_of(seq: int, batch: int, world: int):
    """(rank, kv_slot) of sequence `seq`"""
    per = batch // world
    return seq // per, seq % per


def gather_tokens(local_tokens: torch.Tensor, out: torch.Tensor = None, group=None) -> torch.Tensor:
    """all-gather of this rank's sampled tokens (int32 [batch / world]) -> int32 [batch] in global sequence order.
    On a CUDA tensor the collective is enqueued on the current stream (the engine stream in bench.py): no host sync."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_tokens if out is None else out.copy_(local_tokens)
    if out is None:
        out = torch.empty(local_tokens.numel() * world, dtype=local_tokens.dtype, device=local_tokens.device)
    dist.all_gather_into_tensor(out, local_tokens.contiguous(), group=group)
    return out


def finished_mask(all_tokens: torch.Tensor, already: torch.Tensor) -> torch.Tensor:
    """generation of a sequence ends on token id 0 (transformer.cpp:93); once finished, always finished"""
    return already | (all_tokens == 0)
