"""Device-side run-level symbol generation (count -> prefix scan -> emit) against the oracle and the reference text."""
import numpy as np
import pytest

from oracle import codec_oracle as co
from tests.golden_util import case_names, load_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", [n for n in case_names() if n.startswith("s_") or n == "c1_cif"])
def test_symbols_match_oracle_and_reference_text(name):
    from streamoptima_b200.Encoder import Y_Video_codec
    Y_Video_codec.write_recon_yuv = False
    frames, enc, g = load_case(name)
    F, H, W = frames.shape
    e = dict(enc)
    bs = e["block_size"]
    c = Y_Video_codec(H, W, F, e.pop("block_size"), e.pop("search_range"), e.pop("Qp"), e.pop("intra_dur"), 0,
                      y_only_frame_arr=frames, **e)
    c.encode()
    offsets, symbols, base = c.symbol_streams()
    p = c.encoded_package.packed
    nbx, sub = W // bs, bs // 2
    rng = np.random.default_rng(0)
    nblk = p["split"].shape[1]
    for f in range(F):
        assert offsets[f, -1] == p["qsize"][f]                     # quantized_sized is the frame's symbol count
        for b in rng.choice(nblk, size=min(nblk, 24), replace=False):
            y, x = (b // nbx) * bs, (b % nbx) * bs
            if p["split"][f, b] == 0:
                blocks = [p["levels"][f, y:y + bs, x:x + bs]]
            else:
                blocks = [p["levels"][f, y + (k // 2) * sub:y + (k // 2) * sub + sub, x + (k % 2) * sub:x + (k % 2) * sub + sub]
                          for k in range(4)]
            for k, blkv in enumerate(blocks):
                s0, s1 = int(base[f]) + int(offsets[f, 4 * b + k]), int(base[f]) + int(offsets[f, 4 * b + k + 1])
                assert symbols[s0:s1].tolist() == co.rle_symbols(blkv)
    lines = c.residual_lines_from_symbols()
    assert "".join(l + "\n" for l in lines) == g["res_text"]
