"""`load` (ramba/ramba.py:8930-8945, ramba/fileio.py): distributed reads of file types that can be read in parts (every rank
reads only its block), whole-file reads for images, handler selection by extension / explicit type."""
import numpy as onp
import pytest


def _write_files(tmp_path):
    rng = onp.random.RandomState(11)
    a2 = rng.randint(-100, 100, size=(37, 53)).astype(onp.float32)
    a1 = onp.arange(1000, dtype=onp.int64) * 3 - 7
    p2, p1 = str(tmp_path / "a2.npy"), str(tmp_path / "a1.npy")
    onp.save(p2, a2)
    onp.save(p1, a1)
    return (p1, a1), (p2, a2)


def test_npy_is_loaded_block_by_block(oracle_engine, tmp_path, monkeypatch):
    import ramba_b200 as rb
    from ramba_b200 import fileio

    (p1, a1), (p2, a2) = _write_files(tmp_path)
    reads = []
    h = fileio.get_load_handler(p2)
    orig = h.read
    monkeypatch.setattr(h, "read", lambda fname, index, **kw: reads.append(index) or orig(fname, index, **kw))
    monkeypatch.setattr(h, "readall", lambda *a, **k: (_ for _ in ()).throw(AssertionError("whole-file read of a distributed type")))
    x = rb.load(p2)
    assert x.shape == a2.shape and x.dtype == a2.dtype and len(reads) == 1
    assert onp.array_equal(x.asarray(), a2)
    y = rb.load(p1, dtype=onp.float64)
    assert y.dtype == onp.float64 and onp.array_equal((y * 0.5).asarray(), a1 * 0.5)
    # loaded arrays are ordinary operands of the fused path
    assert float((x * 2.0 + 1.0).sum()) == float((a2.astype(onp.float64) * 2.0 + 1.0).sum())


def test_local_flag_and_explicit_type(oracle_engine, tmp_path):
    import ramba_b200 as rb

    (p1, a1), _ = _write_files(tmp_path)
    other = str(tmp_path / "noext")
    import shutil

    shutil.copy(p1, other)
    assert onp.array_equal(rb.load(other, ftype="npy").asarray(), a1)
    assert onp.array_equal(rb.load(p1, local=True).asarray(), a1)
    with pytest.raises(ValueError):
        rb.load(str(tmp_path / "x.unknowntype"))


def test_images_come_channels_first(oracle_engine, tmp_path):
    PIL = pytest.importorskip("PIL.Image")
    import ramba_b200 as rb

    rng = onp.random.RandomState(3)
    rgb = rng.randint(0, 256, size=(20, 31, 3)).astype(onp.uint8)
    gray = rng.randint(0, 256, size=(17, 9)).astype(onp.uint8)
    p_rgb, p_gray = str(tmp_path / "c.png"), str(tmp_path / "g.png")
    PIL.fromarray(rgb).save(p_rgb)
    PIL.fromarray(gray).save(p_gray)
    x = rb.load(p_rgb)
    assert x.shape == (3, 20, 31) and x.dtype == onp.uint8
    assert onp.array_equal(x.asarray(), onp.transpose(rgb, (2, 0, 1)))
    assert onp.array_equal(rb.load(p_gray).asarray(), gray)
    assert onp.array_equal((rb.load(p_gray).astype(onp.float32) * 2.0).asarray(), gray.astype(onp.float32) * 2.0)
