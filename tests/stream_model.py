"""Pure-Python restatement of STFTStreamer.ProcessChunk's buffer handling (analyzers/spectral.go:323-374): which
sample ranges become frames, and what stays buffered.  Used by the CPU and GPU streaming tests."""


def frame_starts(chunk_sizes, win, hop):
    """-> (per chunk: list of absolute start positions of the frames it yields, samples left in the buffer)."""
    buf_start, buf_len, out = 0, 0, []
    for c in chunk_sizes:
        starts = []
        if c > 0:  # :324-326 an empty chunk returns before touching the buffer
            buf_len += c
            while buf_len >= win:  # :333
                starts.append(buf_start)
                if hop >= buf_len:  # :364-366 the buffer is emptied, the rest of the hop is NOT skipped later
                    buf_start += buf_len
                    buf_len = 0
                else:  # :367-369
                    buf_start += hop
                    buf_len -= hop
        out.append(starts)
    return out, buf_len
