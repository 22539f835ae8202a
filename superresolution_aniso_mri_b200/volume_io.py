"""Minimal volume file IO for the CLI where SimpleITK is not installed: NIfTI-1 (.nii / .nii.gz) and MetaImage
(.mha, .mhd + raw) read / write with the spacing bookkeeping of ``generate_hr_volumes.py:104-127,136-156,177-182``.

The reference does all file IO through SimpleITK (``sitk.ReadImage`` / ``GetArrayFromImage`` / ``SetSpacing`` /
``WriteImage``); there is no arithmetic in it.  This module is a fallback, not a re-implementation of ITK: it reads the
voxel array as SimpleITK would hand it over (``[z, y, x]`` or ``[t, z, y, x]``, intensity scaling applied), keeps the
geometry of the input header, and writes the synthesized volume back with the new slice count and z spacing.  PARITY
UNPINNED against SimpleITK (absent from this image): pinned here are the format's own invariants (round trips, header
fields, both byte orders, gzip) -- ``tests/test_volume_io.py``.

NIfTI-1 header fields used (byte offsets): sizeof_hdr 0 (=348), dim[8] 40 (int16), datatype 70, bitpix 72, pixdim[8] 76
(float32), vox_offset 108, scl_slope 112, scl_inter 116, qform_code 252, sform_code 254, srow_x/y/z 280/296/312, magic 344.
"""
from __future__ import annotations

import gzip
import os
import zlib
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

_NIFTI_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
                 768: np.uint32, 1024: np.int64, 1280: np.uint64}
_MET_TYPES = {"MET_UCHAR": np.uint8, "MET_CHAR": np.int8, "MET_USHORT": np.uint16, "MET_SHORT": np.int16,
              "MET_UINT": np.uint32, "MET_INT": np.int32, "MET_ULONG": np.uint64, "MET_LONG": np.int64,
              "MET_FLOAT": np.float32, "MET_DOUBLE": np.float64}


@dataclass
class Volume:
    """What the CLI needs of a ``sitk.Image``: the array as ``GetArrayFromImage`` returns it and the geometry."""
    array: np.ndarray                                  # [z,y,x] or [t,z,y,x]
    spacing: Tuple[float, ...]                         # (x, y, z[, t]) like sitk.Image.GetSpacing()
    origin: Tuple[float, ...] = ()
    direction: Tuple[float, ...] = ()
    fmt: str = "nifti"                                 # "nifti" | "mha"
    header: Optional[bytes] = field(default=None, repr=False)   # NIfTI: the 348 header bytes of the source file
    byteorder: str = "<"

    def GetSpacing(self):
        return tuple(self.spacing)

    def GetSize(self):
        return tuple(int(n) for n in self.array.shape[::-1])


# ------------------------------------------------------------------------------------------------------------ NIfTI-1
def _open(path: str, mode: str):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def read_nifti(path) -> Volume:
    with _open(str(path), "rb") as f:
        raw = f.read()
    if len(raw) < 348:
        raise ValueError("%s: not a NIfTI-1 file (shorter than its header)" % path)
    bo = "<" if int(np.frombuffer(raw, "<i4", 1, 0)[0]) == 348 else ">"
    if int(np.frombuffer(raw, bo + "i4", 1, 0)[0]) != 348 or raw[344:347] not in (b"n+1", b"ni1"):
        raise ValueError("%s: not a NIfTI-1 file" % path)
    if raw[344:347] == b"ni1":
        raise ValueError("%s: NIfTI pairs (.hdr/.img) are not supported, single-file .nii only" % path)
    dim = np.frombuffer(raw, bo + "i2", 8, 40)
    nd = int(dim[0])
    if not 3 <= nd <= 4:
        raise ValueError("%s: %d-dimensional image (3-D volumes and 4-D series only)" % (path, nd))
    shape_xyz = [int(v) for v in dim[1:nd + 1]]
    code = int(np.frombuffer(raw, bo + "i2", 1, 70)[0])
    if code not in _NIFTI_DTYPES:
        raise ValueError("%s: NIfTI datatype %d not supported" % (path, code))
    dt = np.dtype(_NIFTI_DTYPES[code]).newbyteorder(bo)
    pixdim = np.frombuffer(raw, bo + "f4", 8, 76)
    vox = int(np.frombuffer(raw, bo + "f4", 1, 108)[0])
    slope, inter = (float(v) for v in np.frombuffer(raw, bo + "f4", 2, 112))
    count = int(np.prod(shape_xyz))
    data = np.frombuffer(raw, dt, count, max(vox, 352)).reshape(shape_xyz[::-1])     # x fastest -> [.., z, y, x]
    data = data.astype(dt.newbyteorder("="))
    if slope not in (0.0, 1.0) or (slope != 0.0 and inter != 0.0):
        data = (data.astype(np.float64) * slope + inter).astype(np.float32)
    srow = np.frombuffer(raw, bo + "f4", 12, 280).reshape(3, 4)
    sform = int(np.frombuffer(raw, bo + "i2", 1, 254)[0])
    origin = tuple(float(v) for v in srow[:, 3]) if sform > 0 else tuple(float(v) for v in np.frombuffer(raw, bo + "f4", 3, 268))
    sp = np.abs(pixdim[1:nd + 1]).astype(np.float64)
    direction = ()
    if sform > 0 and np.all(sp[:3] > 0):
        direction = tuple(float(v) for v in (srow[:, :3] / sp[:3]).reshape(-1))
    return Volume(array=data, spacing=tuple(float(v) for v in sp), origin=origin, direction=direction, fmt="nifti",
                  header=bytes(raw[:348]), byteorder=bo)


def write_nifti(path, array: np.ndarray, like: Optional[Volume] = None, spacing: Optional[Sequence[float]] = None) -> None:
    """Write ``array`` ([z,y,x] / [t,z,y,x]) as float32 NIfTI-1.  Geometry comes from ``like`` (the input volume) with the
    new ``spacing``: dim / pixdim are updated and the affine's column of every axis is rescaled by new / old spacing, so the
    volume keeps its origin and orientation (what SetOrigin / SetDirection / SetSpacing do in the reference, :121-126)."""
    arr = np.ascontiguousarray(array, dtype=np.float32)
    nd = arr.ndim
    if not 3 <= nd <= 4:
        raise ValueError("write_nifti: 3-D or 4-D arrays only")
    hdr = bytearray(like.header) if (like is not None and like.header is not None) else bytearray(348)
    bo = like.byteorder if (like is not None and like.header is not None) else "<"

    def put(fmt, off, vals):
        v = np.asarray(vals, dtype=np.dtype(fmt).newbyteorder(bo))
        hdr[off:off + v.nbytes] = v.tobytes()

    old_sp = list(like.spacing) if like is not None else [1.0] * nd
    new_sp = [float(v) for v in (spacing if spacing is not None else old_sp)]
    while len(new_sp) < nd:
        new_sp.append(1.0)
    put("i4", 0, [348])
    dim = [nd] + list(arr.shape[::-1]) + [1] * (7 - nd)
    put("i2", 40, dim)
    put("i2", 70, [16])
    put("i2", 72, [32])
    pixdim = np.frombuffer(bytes(hdr), np.dtype("f4").newbyteorder(bo), 8, 76).copy()
    if pixdim[0] not in (-1.0, 1.0):
        pixdim[0] = 1.0
    pixdim[1:nd + 1] = new_sp[:nd]
    put("f4", 76, pixdim)
    put("f4", 108, [352.0])
    put("f4", 112, [1.0, 0.0])                       # scl_slope, scl_inter: values are stored as they are
    if like is not None and like.header is not None:
        srow = np.frombuffer(bytes(hdr), np.dtype("f4").newbyteorder(bo), 12, 280).reshape(3, 4).copy()
        for ax in range(3):
            if ax < len(old_sp) and old_sp[ax] > 0:
                srow[:, ax] *= new_sp[ax] / old_sp[ax]
        put("f4", 280, srow.reshape(-1))
    else:
        srow = np.zeros((3, 4), np.float32)
        srow[0, 0], srow[1, 1], srow[2, 2] = new_sp[0], new_sp[1], new_sp[2]
        put("f4", 280, srow.reshape(-1))
        put("i2", 252, [0, 2])                       # qform_code 0, sform_code 2 (aligned)
    hdr[344:348] = b"n+1\0"
    payload = bytes(hdr) + b"\0\0\0\0" + arr.astype(np.dtype("f4").newbyteorder(bo)).tobytes()
    with _open(str(path), "wb") as f:
        f.write(payload)


# ---------------------------------------------------------------------------------------------------------- MetaImage
def read_mha(path) -> Volume:
    path = str(path)
    with open(path, "rb") as f:
        raw = f.read()
    fields, pos = {}, 0
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if "=" not in line:
            continue
        k, v = (s.strip() for s in line.split("=", 1))
        fields[k] = v
        if k == "ElementDataFile":
            break
    nd = int(fields["NDims"])
    if not 3 <= nd <= 4:
        raise ValueError("%s: NDims = %d (3-D volumes and 4-D series only)" % (path, nd))
    size = [int(v) for v in fields["DimSize"].split()]
    dt = np.dtype(_MET_TYPES[fields["ElementType"]])
    msb = fields.get("BinaryDataByteOrderMSB", fields.get("ElementByteOrderMSB", "False")).lower() == "true"
    dt = dt.newbyteorder(">" if msb else "<")
    if fields["ElementDataFile"] == "LOCAL":
        blob = raw[pos:]
    else:
        with open(os.path.join(os.path.dirname(path), fields["ElementDataFile"]), "rb") as f:
            blob = f.read()
    if fields.get("CompressedData", "False").lower() == "true":
        blob = zlib.decompress(blob)
    data = np.frombuffer(blob, dt, int(np.prod(size))).reshape(size[::-1]).astype(dt.newbyteorder("="))
    sp = tuple(float(v) for v in fields.get("ElementSpacing", fields.get("ElementSize", " ".join(["1"] * nd))).split())
    origin = tuple(float(v) for v in fields.get("Offset", fields.get("Position", " ".join(["0"] * nd))).split())
    direction = tuple(float(v) for v in fields.get("TransformMatrix", "").split())
    return Volume(array=data, spacing=sp, origin=origin, direction=direction, fmt="mha")


def write_mha(path, array: np.ndarray, like: Optional[Volume] = None, spacing: Optional[Sequence[float]] = None) -> None:
    arr = np.ascontiguousarray(array, dtype=np.float32)
    nd = arr.ndim
    sp = [float(v) for v in (spacing if spacing is not None else (like.spacing if like is not None else [1.0] * nd))]
    while len(sp) < nd:
        sp.append(1.0)
    origin = list(like.origin) if (like is not None and len(like.origin) == nd) else [0.0] * nd
    direction = list(like.direction) if (like is not None and len(like.direction) == nd * nd) else \
        list(np.eye(nd).reshape(-1))
    lines = ["ObjectType = Image", "NDims = %d" % nd, "BinaryData = True", "BinaryDataByteOrderMSB = False",
             "CompressedData = False", "TransformMatrix = " + " ".join("%.17g" % v for v in direction),
             "Offset = " + " ".join("%.17g" % v for v in origin), "CenterOfRotation = " + " ".join(["0"] * nd),
             "ElementSpacing = " + " ".join("%.17g" % v for v in sp[:nd]),
             "DimSize = " + " ".join(str(int(n)) for n in arr.shape[::-1]), "ElementType = MET_FLOAT"]
    path = str(path)
    if path.endswith(".mhd"):
        rawname = os.path.basename(path)[:-4] + ".raw"
        lines.append("ElementDataFile = " + rawname)
        with open(path, "wb") as f:
            f.write(("\n".join(lines) + "\n").encode("ascii"))
        with open(os.path.join(os.path.dirname(path), rawname), "wb") as f:
            f.write(arr.astype("<f4").tobytes())
    else:
        lines.append("ElementDataFile = LOCAL")
        with open(path, "wb") as f:
            f.write(("\n".join(lines) + "\n").encode("ascii"))
            f.write(arr.astype("<f4").tobytes())


# ------------------------------------------------------------------------------------------------------------ generic
def read_volume(path) -> Volume:
    p = str(path).lower()
    if p.endswith(".nii") or p.endswith(".nii.gz"):
        return read_nifti(path)
    if p.endswith(".mha") or p.endswith(".mhd"):
        return read_mha(path)
    raise ValueError("%s: unknown volume format (nii, nii.gz, mha, mhd)" % path)


def write_volume(path, array: np.ndarray, like: Optional[Volume] = None, spacing: Optional[Sequence[float]] = None) -> None:
    p = str(path).lower()
    if p.endswith(".nii") or p.endswith(".nii.gz"):
        if like is not None and like.fmt != "nifti":       # geometry from a MetaImage source: spacing only
            return write_nifti(path, array, like=None, spacing=spacing if spacing is not None else like.spacing)
        return write_nifti(path, array, like=like, spacing=spacing)
    if p.endswith(".mha") or p.endswith(".mhd"):
        return write_mha(path, array, like=like, spacing=spacing)
    raise ValueError("%s: unknown volume format (nii, nii.gz, mha, mhd)" % path)
