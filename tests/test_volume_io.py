"""CPU tier: the fallback volume IO (NIfTI-1, MetaImage) of the CLI -- format invariants only (round trips, header fields,
byte orders, gzip, intensity scaling, spacing bookkeeping of generate_hr_volumes.py:177-182).  Parity with SimpleITK is
UNPINNED (not installed here)."""
import gzip
import zlib

import numpy as np
import pytest

from superresolution_aniso_mri_b200 import volume_io as V


def _nifti_bytes(arr_zyx, bo="<", code=4, pixdim=(1.0, 1.4, 1.4, 10.0), slope=0.0, inter=0.0, sform=1, srow=None):
    hdr = bytearray(348)

    def put(fmt, off, vals):
        v = np.asarray(vals, dtype=np.dtype(fmt).newbyteorder(bo))
        hdr[off:off + v.nbytes] = v.tobytes()

    nd = arr_zyx.ndim
    put("i4", 0, [348])
    put("i2", 40, [nd] + list(arr_zyx.shape[::-1]) + [1] * (7 - nd))
    put("i2", 70, [code])
    put("i2", 72, [arr_zyx.dtype.itemsize * 8])
    put("f4", 76, list(pixdim) + [0.0] * (8 - len(pixdim)))
    put("f4", 108, [352.0])
    put("f4", 112, [slope, inter])
    put("i2", 252, [0, sform])
    if srow is None:
        srow = np.array([[-pixdim[1], 0, 0, 90.0], [0, -pixdim[2], 0, 120.0], [0, 0, pixdim[3], -30.0]], np.float32)
    put("f4", 280, np.asarray(srow, np.float32).reshape(-1))
    hdr[344:348] = b"n+1\0"
    return bytes(hdr) + b"\0\0\0\0" + arr_zyx.astype(arr_zyx.dtype.newbyteorder(bo)).tobytes()


@pytest.mark.parametrize("bo", ["<", ">"])
@pytest.mark.parametrize("gz", [False, True])
def test_read_nifti_int16_both_byte_orders(tmp_path, bo, gz):
    rs = np.random.RandomState(1)
    arr = rs.randint(-500, 3000, size=(10, 12, 14)).astype(np.int16)           # [z, y, x]
    path = tmp_path / ("vol.nii.gz" if gz else "vol.nii")
    data = _nifti_bytes(arr, bo=bo)
    path.write_bytes(gzip.compress(data) if gz else data)
    v = V.read_volume(path)
    assert v.array.dtype == np.int16 and v.array.shape == (10, 12, 14)
    np.testing.assert_array_equal(v.array, arr)
    assert v.GetSize() == (14, 12, 10)
    np.testing.assert_allclose(v.GetSpacing(), (1.4, 1.4, 10.0), rtol=1e-6)
    np.testing.assert_allclose(v.origin, (90.0, 120.0, -30.0))
    np.testing.assert_allclose(np.array(v.direction).reshape(3, 3), np.diag([-1.0, -1.0, 1.0]), atol=1e-6)


def test_read_nifti_applies_intensity_scaling(tmp_path):
    arr = np.arange(2 * 3 * 4, dtype=np.uint8).reshape(2, 3, 4)
    path = tmp_path / "s.nii"
    path.write_bytes(_nifti_bytes(arr, code=2, slope=0.5, inter=-3.0))
    v = V.read_nifti(path)
    assert v.array.dtype == np.float32
    np.testing.assert_allclose(v.array, arr.astype(np.float64) * 0.5 - 3.0)
    path.write_bytes(_nifti_bytes(arr, code=2, slope=1.0, inter=0.0))
    assert V.read_nifti(path).array.dtype == np.uint8                          # identity scaling keeps the stored type


@pytest.mark.parametrize("suffix", [".nii", ".nii.gz"])
def test_hr_volume_written_with_new_slice_count_and_z_spacing(tmp_path, suffix):
    """The CLI's bookkeeping (generate_hr_volumes.py:177-182): Z -> (Z-1)(ni+1)+1 slices, z spacing / (ni+1), x/y spacing,
    origin and orientation kept."""
    rs = np.random.RandomState(2)
    arr = rs.randint(0, 2000, size=(10, 16, 18)).astype(np.int16)
    rot = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]]) @ np.diag([1.25, 1.25, 8.0])
    srow = np.concatenate([rot, [[5.0], [6.0], [7.0]]], axis=1)
    src = tmp_path / "lr.nii"
    src.write_bytes(_nifti_bytes(arr, pixdim=(-1.0, 1.25, 1.25, 8.0), srow=srow))
    lr = V.read_volume(src)
    ni = 6
    hr = rs.rand((10 - 1) * (ni + 1) + 1, 16, 18).astype(np.float32)
    new_spacing = np.asarray(lr.GetSpacing()[:2] + (lr.GetSpacing()[-1] / (ni + 1),)).astype(np.float64)
    dst = tmp_path / ("hr" + suffix)
    V.write_volume(dst, hr, like=lr, spacing=new_spacing)
    back = V.read_volume(dst)
    assert back.array.dtype == np.float32 and back.array.shape == hr.shape
    np.testing.assert_array_equal(back.array, hr)
    np.testing.assert_allclose(back.GetSpacing(), (1.25, 1.25, 8.0 / 7), rtol=1e-6)
    np.testing.assert_allclose(back.origin, (5.0, 6.0, 7.0))
    np.testing.assert_allclose(np.array(back.direction).reshape(3, 3), np.array(lr.direction).reshape(3, 3), atol=1e-6)
    raw = dst.read_bytes() if suffix == ".nii" else gzip.decompress(dst.read_bytes())
    assert float(np.frombuffer(raw, "<f4", 1, 76)[0]) == -1.0                  # qfac of the source header survives
    assert len(raw) == 352 + hr.size * 4


def test_nifti_4d_series_round_trip(tmp_path):
    arr = np.random.RandomState(3).rand(3, 5, 6, 7).astype(np.float32)         # [t, z, y, x]
    p = tmp_path / "cine.nii.gz"
    V.write_nifti(p, arr, spacing=(1.0, 1.0, 5.0, 1.0))
    v = V.read_nifti(p)
    np.testing.assert_array_equal(v.array, arr)
    assert v.GetSize() == (7, 6, 5, 3) and v.GetSpacing() == (1.0, 1.0, 5.0, 1.0)


def test_rejects_non_nifti_and_unknown_formats(tmp_path):
    p = tmp_path / "x.nii"
    p.write_bytes(b"\0" * 400)
    with pytest.raises(ValueError):
        V.read_nifti(p)
    with pytest.raises(ValueError):
        V.read_volume(tmp_path / "x.dcm")
    with pytest.raises(ValueError):
        V.write_volume(tmp_path / "x.png", np.zeros((2, 2, 2), np.float32))


@pytest.mark.parametrize("suffix", [".mha", ".mhd"])
def test_metaimage_round_trip_and_spacing(tmp_path, suffix):
    arr = np.random.RandomState(4).rand(6, 9, 11).astype(np.float32)
    p = tmp_path / ("a" + suffix)
    V.write_volume(p, arr, spacing=(0.8, 0.8, 6.0))
    v = V.read_volume(p)
    np.testing.assert_array_equal(v.array, arr)
    assert v.GetSpacing() == (0.8, 0.8, 6.0) and v.GetSize() == (11, 9, 6) and v.fmt == "mha"
    hr = np.random.RandomState(5).rand(31, 9, 11).astype(np.float32)
    q = tmp_path / ("b" + suffix)
    V.write_volume(q, hr, like=v, spacing=(0.8, 0.8, 1.0))
    w = V.read_volume(q)
    np.testing.assert_array_equal(w.array, hr)
    assert w.GetSpacing() == (0.8, 0.8, 1.0) and w.origin == v.origin and w.direction == v.direction


def test_metaimage_compressed_short_big_endian(tmp_path):
    arr = np.random.RandomState(6).randint(-100, 100, size=(4, 5, 6)).astype(np.int16)
    hdr = ("ObjectType = Image\nNDims = 3\nBinaryData = True\nBinaryDataByteOrderMSB = True\nCompressedData = True\n"
           "TransformMatrix = 1 0 0 0 1 0 0 0 1\nOffset = 1 2 3\nElementSpacing = 1.5 1.5 7\nDimSize = 6 5 4\n"
           "ElementType = MET_SHORT\nElementDataFile = LOCAL\n").encode("ascii")
    p = tmp_path / "c.mha"
    p.write_bytes(hdr + zlib.compress(arr.astype(">i2").tobytes()))
    v = V.read_mha(p)
    np.testing.assert_array_equal(v.array, arr)
    assert v.spacing == (1.5, 1.5, 7.0) and v.origin == (1.0, 2.0, 3.0)


def test_cli_file_flow_without_simpleitk(tmp_path, monkeypatch):
    """generate_hr_volumes.load_images -> main -> save_images on NIfTI / MetaImage files with the synthesis replaced by a
    stand-in (the kernels need a GPU): slice count, z spacing and geometry of the written files follow :159-183."""
    import torch
    import generate_hr_volumes as ghv
    if ghv.sitk is not None:
        pytest.skip("SimpleITK is installed: the CLI uses it, as the reference does")
    rs = np.random.RandomState(7)
    lr = (rs.rand(5, 8, 9) * 900).astype(np.int16)                      # needs normalisation, like a real scan
    (tmp_path / "in").mkdir()
    (tmp_path / "out").mkdir()
    (tmp_path / "in" / "pat01.nii.gz").write_bytes(gzip.compress(_nifti_bytes(lr, pixdim=(1.0, 1.5, 1.5, 9.0))))
    calls = []

    def fake_csv(trainer, images, alpha_range, use_original=False, labels=None):
        z = images.shape[0]
        a = len(alpha_range)
        assert use_original and images.dim() == 4 and float(images.min()) >= 0 and float(images.max()) <= 1
        calls.append((z, a))
        out = torch.zeros((z - 1) * (a + 1) + 1, images.shape[2], images.shape[3])
        out[::a + 1] = images[:, 0]
        return {"upsampled_image": out}

    monkeypatch.setattr(ghv, "create_super_volume", fake_csv)
    images = ghv.load_images(tmp_path / "in")
    assert len(images) == 1 and isinstance(images[0][1], V.Volume)

    class Args:
        num_interpolations = 2

    res = ghv.main(Args, None, images, tmp_path / "out")
    ghv.save_images(res)
    assert calls == [(5, 2)]
    hr = V.read_volume(tmp_path / "out" / "pat01.nii.gz")
    assert hr.array.shape == (13, 8, 9) and hr.array.dtype == np.float32
    np.testing.assert_allclose(hr.GetSpacing(), (1.5, 1.5, 3.0), rtol=1e-6)
    np.testing.assert_allclose(hr.origin, (90.0, 120.0, -30.0))
    want = ghv.normalize_img(lr.astype(np.float32)).astype(np.float32)     # array_to_torch: float64 math, one fp32 rounding
    np.testing.assert_array_equal(hr.array[::3], want)                  # the kept slices are the normalised inputs
    raw = gzip.decompress((tmp_path / "out" / "pat01.nii.gz").read_bytes())
    srow = np.frombuffer(raw, "<f4", 12, 280).reshape(3, 4)
    np.testing.assert_allclose(srow[:, :3], np.diag([-1.5, -1.5, 3.0]), rtol=1e-6)
