"""Frame-range sharding of one stream over ranks/GPUs.

A FLAC frame depends only on its own block of PCM and its frame number, so a stream is split
into contiguous, block-aligned PCM ranges; every rank encodes its range with
`first_frame_number` set to the index of its first block, and the host concatenates the
returned frame bytes in rank order.  No data-path collective is involved; the only exchange
is the host-side gather of (bytes, per-frame sizes), from which the reference's
(offset, pcm_frames) list (src/encoders/flac.c:249-253) and STREAMINFO's min/max frame size
(flac.c:259-265) follow by a prefix sum.
"""


def frame_ranges(total_pcm_frames, block_size, world):
    """[(pcm_frame_offset, n_pcm_frames, first_frame_number)] per rank; only the last
    non-empty range can end in a short block"""
    n_blocks = (total_pcm_frames + block_size - 1) // block_size
    base, extra = divmod(n_blocks, world)
    out = []
    blk = 0
    for r in range(world):
        nb = base + (1 if r < extra else 0)
        start = blk * block_size
        end = min((blk + nb) * block_size, total_pcm_frames)
        out.append((min(start, total_pcm_frames), max(0, end - start), blk))
        blk += nb
    return out


def merge_frame_tables(per_rank_frame_bytes, per_rank_frame_pcm):
    """rank-ordered per-frame sizes -> ([(byte_offset, pcm_frames)], min_frame, max_frame, total_bytes)"""
    offsets = []
    pos = 0
    mn, mx = 0xFFFFFF, 0
    for sizes, pcms in zip(per_rank_frame_bytes, per_rank_frame_pcm):
        for s, n in zip(sizes, pcms):
            offsets.append((pos, int(n)))
            pos += int(s)
            mn = min(mn, int(s))
            mx = max(mx, int(s))
    return offsets, mn, mx, pos
