"""LimitBuffIndexCopyer (upstream copyer.py).

The reference used it to bound the staging memory of its CPU<->CUDA row copies.  The B200 path moves rows with
zero-copy kernels and needs no staging at all; the class is kept so code that constructs or calls it keeps working.
It handles same-device and host<->device tensors with plain index_select / index_copy_ in bounded pieces.
"""
import torch


class LimitBuffIndexCopyer(object):

    def __init__(self, size: int) -> None:
        self._buff_size = size

    @torch.no_grad()
    def index_copy(self, dim: int, src_index: torch.LongTensor, tgt_index: torch.LongTensor, src: torch.Tensor,
                   tgt: torch.Tensor):
        """tgt.index_copy_(dim, tgt_index, src.index_select(dim, src_index)) in pieces of at most `size` rows."""
        dim_size = src_index.numel()
        src_index = src_index.to(src.device)
        for begin_pos in range(0, dim_size, self._buff_size):
            cur_len = min(self._buff_size, dim_size - begin_pos)
            src_idx_piece = src_index.narrow(0, begin_pos, cur_len)
            if src.device.type == 'cpu' and tgt.device.type == 'cuda':
                cpu_part = src.index_select(dim, src_idx_piece).pin_memory()
                tmp_buffer = torch.empty_like(cpu_part, device=tgt.device)
                tmp_buffer.copy_(cpu_part)
            else:
                tmp_buffer = src.index_select(dim, src_idx_piece).to(tgt.device)
            tgt_idx_piece = tgt_index.narrow(0, begin_pos, cur_len)
            tgt.index_copy_(dim, tgt_idx_piece.to(tgt.device), tmp_buffer)
