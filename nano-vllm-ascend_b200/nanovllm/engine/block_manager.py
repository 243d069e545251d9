"""Paged-KV block allocator with a chained-hash prefix cache.

Host bookkeeping only: it produces the block tables the CUDA kernels consume. Observable
behaviour follows the reference (engine/block_manager.py): a FULL block is identified by
xxh64(previous block's hash || its token ids as int64 bytes) (:38-44); a prompt re-uses cached
blocks up to its first miss and takes fresh blocks after it, registering every full block's
hash on the way (:62-88); after a token is appended, ``may_append`` takes a new block when the
sequence has just spilled into one and registers the hash of a block that has just filled
(:102-118). Attribute names (blocks, hash_to_block_id, free_block_ids, used_block_ids) are the
reference's so its unit tests (ut/test_block_manager.py) can drive this class.
"""
from __future__ import annotations

from array import array
from collections import deque

import xxhash

from .sequence import Sequence

NO_HASH = -1


class Block:
    __slots__ = ("block_id", "ref_count", "hash", "token_ids")

    def __init__(self, block_id: int):
        self.block_id = block_id
        self.ref_count = 0
        self.hash = NO_HASH
        self.token_ids: list[int] = []

    def update(self, h: int, token_ids: list[int]) -> None:
        self.hash, self.token_ids = h, token_ids

    def reset(self) -> None:
        self.ref_count, self.hash, self.token_ids = 1, NO_HASH, []


class BlockManager:

    def __init__(self, num_blocks: int, block_size: int, non_cache_token_ids=None):
        self.block_size = block_size
        self.blocks = [Block(i) for i in range(num_blocks)]
        self.hash_to_block_id: dict[int, int] = {}
        self.free_block_ids: deque[int] = deque(range(num_blocks))
        self.used_block_ids: set[int] = set()
        self.non_cache_token_ids = frozenset(non_cache_token_ids or ())

    # ------------------------------------------------------------------ hashing
    @classmethod
    def compute_hash(cls, token_ids: list[int], prefix: int = NO_HASH) -> int:
        digest = xxhash.xxh64()
        if prefix != NO_HASH:
            digest.update(prefix.to_bytes(8, "little"))
        digest.update(array("q", token_ids).tobytes())   # == np.array(ids, int64).tobytes()
        return digest.intdigest()

    def _chain(self, seq: Sequence) -> list[int]:
        """Hash of every block of the sequence (NO_HASH for a partial last block)."""
        hashes, h = [], NO_HASH
        for i in range(seq.num_blocks):
            tokens = seq.block(i)
            h = self.compute_hash(tokens, h) if len(tokens) == self.block_size else NO_HASH
            hashes.append(h)
        return hashes

    # ------------------------------------------------------------------ pool
    def _take(self, block_id: int) -> Block:
        block = self.blocks[block_id]
        assert block.ref_count == 0
        block.reset()
        self.free_block_ids.remove(block_id)
        self.used_block_ids.add(block_id)
        return block

    def _give_back(self, block_id: int) -> None:
        assert self.blocks[block_id].ref_count == 0
        self.used_block_ids.remove(block_id)
        self.free_block_ids.append(block_id)

    # ------------------------------------------------------------------ prompt admission
    def can_allocate(self, seq: Sequence) -> bool:
        return len(self.free_block_ids) >= seq.num_blocks

    def allocate(self, seq: Sequence) -> None:
        assert not seq.block_table
        # a re-admitted (preempted) sequence starts from scratch: the reference never resets this
        # counter (block_manager.py:62-88) and ignores it in its runner; here prepare_prefill and the
        # chunked scheduler consume it, so a stale value would skip tokens whose K/V are gone
        seq.num_cached_tokens = 0
        hashes = self._chain(seq)
        # length of the cached prefix: leading blocks whose hash is known AND whose stored
        # tokens really are these tokens (hash collisions do not count as hits)
        hits = 0
        for i, h in enumerate(hashes):
            tokens = seq.block(i)
            cached = self.hash_to_block_id.get(h, -1)
            if cached == -1 or self.blocks[cached].token_ids != tokens:
                break
            if self.non_cache_token_ids and not self.non_cache_token_ids.isdisjoint(tokens):
                break
            hits += 1
        for i, h in enumerate(hashes):
            if i < hits:
                block_id = self.hash_to_block_id[h]
                seq.num_cached_tokens += self.block_size
                if block_id in self.used_block_ids:
                    block = self.blocks[block_id]
                    block.ref_count += 1
                else:                       # cached content still sits in a freed block
                    block = self._take(block_id)
            else:
                block_id = self.free_block_ids[0]
                block = self._take(block_id)
            if h != NO_HASH:
                block.update(h, seq.block(i))
                self.hash_to_block_id[h] = block_id
            seq.block_table.append(block_id)

    def forget_uncomputed(self, seq: Sequence, num_computed_tokens: int) -> None:
        """Chunked prefill: ``allocate`` registered the hash of every full prompt block up front. If
        the sequence leaves before its prefill completed, blocks past ``num_computed_tokens`` that
        nobody else holds never received their K/V: drop their hashes so that no later prompt can
        take them for cached content."""
        first = num_computed_tokens // self.block_size
        for block_id in seq.block_table[first:]:
            block = self.blocks[block_id]
            if block.ref_count == 1 and block.hash != NO_HASH:
                if self.hash_to_block_id.get(block.hash) == block_id:
                    del self.hash_to_block_id[block.hash]
                block.update(NO_HASH, [])

    def deallocate(self, seq: Sequence) -> None:
        for block_id in reversed(seq.block_table):
            block = self.blocks[block_id]
            block.ref_count -= 1
            if block.ref_count == 0:
                self._give_back(block_id)
        seq.block_table.clear()

    # ------------------------------------------------------------------ decode growth
    def can_append(self, seq: Sequence) -> bool:
        needs_block = len(seq) % self.block_size == 1
        return len(self.free_block_ids) >= int(needs_block)

    def may_append(self, seq: Sequence) -> None:
        table = seq.block_table
        last = self.blocks[table[-1]]
        fill = len(seq) % self.block_size
        if fill == 1:                       # the newest token opened a block
            assert last.hash != NO_HASH
            fresh = self.free_block_ids[0]
            self._take(fresh)
            table.append(fresh)
        elif fill == 0:                     # the newest token completed a block
            assert last.hash == NO_HASH
            tokens = seq.block(seq.num_blocks - 1)
            prev = self.blocks[table[-2]].hash if len(table) > 1 else NO_HASH
            h = self.compute_hash(tokens, prev)
            last.update(h, tokens)
            self.hash_to_block_id[h] = last.block_id
        else:
            assert last.hash == NO_HASH
