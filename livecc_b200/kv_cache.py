"""Paged KV cache: the object that lives in `state['past_key_values']` between `generate()` calls.

Replaces transformers' DynamicCache/DynamicLayer (cache_utils.py:88-163), which re-concatenates the
whole K and V of every layer on every token (`torch.cat`, :119-120). Here a stream owns a list of
64-token pages inside one pool per engine; appending a token writes 2*layers*kv_heads*128 bf16 and
never moves old data. The per-stream rope_delta (kept on the *model* by the reference,
mq2vl.py:923) lives here, so any number of streams can interleave on one engine.

Pool layout (HBM): K and V pools are bf16 [layers, num_pages, kv_heads, 64, 128]; a layer's pool is
contiguous so a (page, head) tile is one 16 KB run — coalesced for the decode kernel and TMA-shaped.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ._cabi import PAGE_SIZE, SC_COUNT, StreamState


class PagePool:
    """Fixed-size pages shared by all streams of one engine. Grows by doubling (rare)."""

    def __init__(self, layers: int, kv_heads: int, device, initial_pages: int = 160):
        self.layers, self.kv_heads, self.device = layers, kv_heads, device
        self.num_pages = 0
        self.k = self.v = None
        self.free: List[int] = []
        self.generation = 0  # bumped whenever the pool storage moves (invalidates captured graphs)
        self._free_buffers: List["StreamBuffers"] = []
        self._grow(initial_pages)

    def _grow(self, new_total: int):
        shape = (self.layers, new_total, self.kv_heads, PAGE_SIZE, 128)
        k = torch.empty(shape, dtype=torch.bfloat16, device=self.device)
        v = torch.empty(shape, dtype=torch.bfloat16, device=self.device)
        if self.k is not None:
            k[:, : self.num_pages].copy_(self.k)
            v[:, : self.num_pages].copy_(self.v)
        self.free.extend(range(self.num_pages, new_total))
        self.k, self.v, self.num_pages = k, v, new_total
        self.generation += 1

    @property
    def layer_stride(self) -> int:
        return self.num_pages * self.kv_heads * PAGE_SIZE * 128

    def alloc(self, n: int) -> List[int]:
        while len(self.free) < n:
            self._grow(max(self.num_pages * 2, self.num_pages + n))
        out = self.free[:n]
        del self.free[:n]
        return out

    def release(self, pages: List[int]):
        self.free.extend(pages)

    # Per-stream device buffers (page table, scalars, id buffer) are recycled through the pool so that their
    # addresses — which are baked into captured decode graphs — stay stable from one stream to the next.
    def acquire_buffers(self) -> "StreamBuffers":
        return self._free_buffers.pop() if self._free_buffers else StreamBuffers(self.device)

    def release_buffers(self, b: "StreamBuffers"):
        self._free_buffers.append(b)

    def bytes_per_token(self) -> int:
        return 2 * self.layers * self.kv_heads * 128 * 2


class StreamBuffers:
    def __init__(self, device):
        self.page_table = torch.zeros(512, dtype=torch.int32, device=device)   # 32k tokens before the first growth
        self.scalars = torch.zeros(SC_COUNT, dtype=torch.int32, device=device)
        self.seq_buf = torch.zeros(32768, dtype=torch.int64, device=device)
        self.generation = 0


class PagedKVCache:
    """One stream's cache: page list, sequence length, rope_delta, device scalars and the id buffer."""

    def __init__(self, pool: PagePool):
        self.pool = pool
        self.pages: List[int] = []
        self.seq_len = 0                       # tokens whose K/V are in the cache
        self.rope_delta: Optional[int] = None  # turn-0 value, never updated (mq2vl.py:1207,1518)
        self._buf = pool.acquire_buffers()
        self._pages_uploaded = 0

    # the buffers live in a recyclable slot; expose them under their old names
    @property
    def page_table(self):
        return self._buf.page_table

    @property
    def scalars(self):
        return self._buf.scalars

    @property
    def seq_buf(self):
        return self._buf.seq_buf

    @property
    def buffers_generation(self):
        return self._buf.generation

    # -- HF Cache surface used by callers (gen/utils.py:3748) -----------------------------------
    def get_seq_length(self, layer_idx: int = 0) -> int:
        return self.seq_len

    def __len__(self):
        return self.pool.layers

    # -- capacity --------------------------------------------------------------------------------
    def ensure_tokens(self, n_tokens: int):
        need = (n_tokens + PAGE_SIZE - 1) // PAGE_SIZE
        if need > len(self.pages):
            self.pages.extend(self.pool.alloc(need - len(self.pages)))
        if need > self.page_table.numel():
            self._buf.page_table = torch.zeros(max(need, 2 * self.page_table.numel()), dtype=torch.int32,
                                               device=self.pool.device)
            self._pages_uploaded = 0
            self._buf.generation += 1
        if self._pages_uploaded < len(self.pages):
            host = torch.tensor(self.pages[self._pages_uploaded:], dtype=torch.int32)
            self.page_table[self._pages_uploaded:len(self.pages)].copy_(host, non_blocking=False)
            self._pages_uploaded = len(self.pages)

    def ensure_seq_capacity(self, n_ids: int):
        if n_ids > self.seq_buf.numel():
            self._buf.seq_buf = torch.zeros(max(n_ids, 2 * self.seq_buf.numel()), dtype=torch.int64,
                                            device=self.pool.device)
            self._buf.generation += 1

    def stream_state(self) -> StreamState:
        p = self.pool
        return StreamState(p.k.data_ptr(), p.v.data_ptr(), p.layer_stride, self.page_table.data_ptr(),
                           self.scalars.data_ptr(), self.seq_buf.data_ptr())

    def graph_key(self):
        return (self.pool.generation, self.buffers_generation, self.page_table.data_ptr(), self.scalars.data_ptr(),
                self.seq_buf.data_ptr())

    def release(self):
        """Return the pages to the pool (end of stream)."""
        if self.pages:
            self.pool.release(self.pages)
            self.pages = []
        if self._buf is not None:
            self.pool.release_buffers(self._buf)
            self._buf = None
        self.seq_len = 0
        self.rope_delta = None
        self._pages_uploaded = 0

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    # -- debugging / parity ------------------------------------------------------------------------
    def gather(self, layer: int):
        """Logical (K, V) of one layer as [kv_heads, seq_len, 128] tensors (test helper)."""
        n = (self.seq_len + PAGE_SIZE - 1) // PAGE_SIZE
        idx = torch.tensor(self.pages[:n], dtype=torch.long, device=self.pool.device)
        k = self.pool.k[layer][idx].permute(1, 0, 2, 3).reshape(self.pool.kv_heads, -1, 128)[:, : self.seq_len]
        v = self.pool.v[layer][idx].permute(1, 0, 2, 3).reshape(self.pool.kv_heads, -1, 128)[:, : self.seq_len]
        return k, v
