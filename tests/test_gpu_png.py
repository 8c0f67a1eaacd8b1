"""GPU PNG writer (lerf_png_encode_stored): the file produced on the device must be a valid PNG that
PIL decodes to exactly the input image, with the checksums zlib computes on the host."""
import io
import struct
import zlib

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lp():
    import lerf_pytorch_b200 as lp
    return lp


def _chunks(data):
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    p, out = 8, []
    while p < len(data):
        n, = struct.unpack(">I", data[p:p + 4])
        typ, body = data[p + 4:p + 8], data[p + 8:p + 8 + n]
        crc, = struct.unpack(">I", data[p + 8 + n:p + 12 + n])
        assert crc == zlib.crc32(typ + body), typ
        out.append((typ, body))
        p += 12 + n
    assert p == len(data)
    return out


# sizes around the 65535-byte stored-block limit, odd shapes, every channel count, one frame of the benched output size
@pytest.mark.parametrize("shape", [(1, 1, 3), (1, 1, 1), (5, 7, 3), (33, 31, 1), (17, 19, 2), (16, 16, 4), (255, 257, 1),
                                   (3, 21844, 3), (5, 13106, 1), (300, 219, 3), (701, 1033, 3), (5424, 8160, 3)])
def test_png_roundtrip_and_checksums(lp, shape):
    H, W, C = shape
    rng = np.random.default_rng(H * 131 + W * 7 + C)
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    d = torch.from_numpy(img).cuda()
    png = lp.encode_png(d if C > 1 else d[:, :, 0])
    assert png.numel() == lp.png_bytes(H, W, C)
    data = png.cpu().numpy().tobytes()
    chunks = _chunks(data)  # every chunk CRC equals zlib.crc32
    assert [t for t, _ in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    raw = zlib.decompress(chunks[1][1])  # verifies the stored-block framing and the Adler-32
    assert len(raw) == H * (1 + W * C)
    got = np.asarray(Image.open(io.BytesIO(data)))
    assert got.shape == ((H, W) if C == 1 else (H, W, C))
    assert np.array_equal(got.reshape(H, W, C), img)


def test_png_of_a_path_result_and_errors(lp, tmp_path):
    import util
    luts = lp.LutSet(lp.load_lut_dict(util.lut_dir("lerf-g")))
    img = torch.from_numpy(util.natural_image(3, 40, 56)).cuda()
    out = lp.LerfSR(luts, 2.5)(img, out_format="u8_hwc")
    assert out.dim() == 3 and out.shape[2] == 3
    path = tmp_path / "x.png"
    n = lp.save_png(out, str(path))
    assert n == lp.png_bytes(*out.shape) and path.stat().st_size == n
    assert np.array_equal(np.asarray(Image.open(str(path))), out.cpu().numpy())
    with pytest.raises(ValueError):
        lp.encode_png(out.float())
    with pytest.raises(ValueError):
        lp.encode_png(out.cpu())
    with pytest.raises(ValueError):
        lp.encode_png(torch.zeros((4, 4, 5), dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        lp.encode_png(out, out=torch.empty(16, dtype=torch.uint8, device="cuda"))
