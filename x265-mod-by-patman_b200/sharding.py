"""Frame sharding across GPUs and the one exchange the path has: reconstructed reference pictures.

Mirrors how the reference parallelises over frames (frame threads gated by recon rows,
source/encoder/frameencoder.cpp:975-978): rank r owns frames {f : f % world == r}; a frame's analysis needs
the reconstructed pictures of its references, which other ranks produced, so each step every rank
contributes the padded recon plane it produced and receives everyone else's (all-gather).  Works on any
torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def frames_for_rank(nframes, rank, world):
    """frame indices owned by `rank` (round-robin, like --frame-threads' in-flight frames)"""
    return list(range(rank, nframes, world))


def owner_of(frame, world):
    return frame % world


def exchange_recon(recon_local, world, group=None, out=None):
    """all-gather one recon plane per rank; returns a (world, plane_elems) tensor whose row r is the plane
    rank r produced.  `out` may be a preallocated flat tensor of world * numel elements."""
    flat = recon_local.reshape(-1)
    if out is None:
        out = torch.empty(world * flat.numel(), dtype=flat.dtype, device=flat.device)
    if world == 1:
        out.copy_(flat)
    else:
        # move raw bytes: pictures are uint8/int16 planes and not every backend has 16-bit integer types
        dist.all_gather_into_tensor(out.view(torch.uint8), flat.contiguous().view(torch.uint8), group=group)
    return out.view(world, flat.numel())


def reference_planes(frame, nrefs, world, gathered_by_step):
    """planes of the `nrefs` previous frames of `frame` from the per-step gather results
    (gathered_by_step[s][r] = recon of frame s * world + r)"""
    refs = []
    for k in range(1, nrefs + 1):
        f = frame - k
        if f < 0:
            break
        refs.append(gathered_by_step[f // world][owner_of(f, world)])
    return refs
