"""Helpers shared by the GPU parity tests (all calls go through the C ABI of libskf.so)."""
import torch



def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_2 / ||b||_2 in float64 (the parity metric of SURVEY §7 'hard parts')."""
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    den = float(b.norm())
    num = float((a - b).norm())
    return num / den if den > 0 else num


from sketchformer_b200.ops import gemm, ptr, stream  # noqa: F401,E402


def to_tf32(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest (ties away, like cvt.rna) fp32 -> tf32, on any device."""
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)
