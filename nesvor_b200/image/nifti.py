"""Minimal NIfTI-1 single-file reader / writer (numpy only).

The reference reads and writes its stacks, slices and volumes through nibabel (nesvor/image/image.py:251-296:
`nib.load`, `img.get_fdata()`, `img.header["pixdim"]`, `img.affine` / `img.get_qform()`, `nib.nifti1.Nifti1Image`,
`set_xyzt_units(2)`, `set_qform(affine, code="aligned")`, `set_sform(affine, code="scanner")`, `nib.save`).  nibabel is
not available in this image, so the handful of calls the path needs is restated here from the published NIfTI-1
standard (nifti1.h: the 348-byte header; nifti1_io.c: `nifti_quatern_to_mat44` / `nifti_mat44_to_quatern`):

* `.nii` and `.nii.gz`, both byte orders (detected from `sizeof_hdr == 348`), magic `n+1`;
* voxel types uint8 / int8 / int16 / uint16 / int32 / uint32 / int64 / uint64 / float32 / float64, `scl_slope` /
  `scl_inter` applied like `get_fdata()` (slope 0 or NaN = no scaling);
* best affine like nibabel's `img.affine`: sform when `sform_code > 0`, else qform when `qform_code > 0`, else the
  pixdim diagonal with the origin at the volume centre;
* written files carry float32 voxels, `xyzt_units = 2` (mm), qform code 2 ("aligned") and sform code 1 ("scanner"),
  `pixdim` = column norms of the affine, `qfac` = sign of its determinant -- what the reference's `save_nii_volume` sets.
Header extensions are skipped on read (voxel data start at `vox_offset`) and never written.
"""
import gzip
import struct
from typing import Dict, Optional, Tuple

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).name: k for k, v in _DTYPES.items()}


def quatern_to_mat44(b: float, c: float, d: float, qx: float, qy: float, qz: float, dx: float, dy: float, dz: float,
                     qfac: float) -> np.ndarray:
    """nifti1_io.c nifti_quatern_to_mat44: unit quaternion (a recovered from b, c, d) + offsets + spacings -> 4x4."""
    a = 1.0 - (b * b + c * c + d * d)
    if a < 1e-7:  # special case: 180 degree rotation
        a = 1.0 / np.sqrt(b * b + c * c + d * d)
        b, c, d = b * a, c * a, d * a
        a = 0.0
    else:
        a = np.sqrt(a)
    xd = dx if dx > 0 else 1.0
    yd = dy if dy > 0 else 1.0
    zd = dz if dz > 0 else 1.0
    if qfac < 0:
        zd = -zd
    R = np.eye(4)
    R[0, :3] = [(a * a + b * b - c * c - d * d) * xd, 2.0 * (b * c - a * d) * yd, 2.0 * (b * d + a * c) * zd]
    R[1, :3] = [2.0 * (b * c + a * d) * xd, (a * a + c * c - b * b - d * d) * yd, 2.0 * (c * d - a * b) * zd]
    R[2, :3] = [2.0 * (b * d - a * c) * xd, 2.0 * (c * d + a * b) * yd, (a * a + d * d - c * c - b * b) * zd]
    R[:3, 3] = [qx, qy, qz]
    return R


def mat44_to_quatern(A: np.ndarray) -> Tuple[float, float, float, float, float, float, float, float, float, float]:
    """nifti1_io.c nifti_mat44_to_quatern: 4x4 -> (b, c, d, qx, qy, qz, dx, dy, dz, qfac).  The 3x3 block is
    orthogonalised by polar decomposition after the column norms have been taken out, as the C code does."""
    A = np.asarray(A, np.float64)
    qx, qy, qz = A[:3, 3]
    M = A[:3, :3].copy()
    dx, dy, dz = np.sqrt((M * M).sum(0))
    if dx == 0:
        M[:, 0], dx = [1, 0, 0], 1.0
    if dy == 0:
        M[:, 1], dy = [0, 1, 0], 1.0
    if dz == 0:
        M[:, 2], dz = [0, 0, 1], 1.0
    M = M / np.array([dx, dy, dz])
    U, _, Vt = np.linalg.svd(M)  # closest orthogonal matrix
    P = U @ Vt
    qfac = 1.0
    if np.linalg.det(P) < 0:
        P[:, 2] = -P[:, 2]
        qfac = -1.0
    r11, r12, r13 = P[0]
    r21, r22, r23 = P[1]
    r31, r32, r33 = P[2]
    a = r11 + r22 + r33 + 1.0
    if a > 0.5:
        a = 0.5 * np.sqrt(a)
        b = 0.25 * (r32 - r23) / a
        c = 0.25 * (r13 - r31) / a
        d = 0.25 * (r21 - r12) / a
    else:
        xd, yd, zd = 1.0 + r11 - (r22 + r33), 1.0 + r22 - (r11 + r33), 1.0 + r33 - (r11 + r22)
        if xd > 1.0:
            b = 0.5 * np.sqrt(xd)
            c = 0.25 * (r12 + r21) / b
            d = 0.25 * (r13 + r31) / b
            a = 0.25 * (r32 - r23) / b
        elif yd > 1.0:
            c = 0.5 * np.sqrt(yd)
            b = 0.25 * (r12 + r21) / c
            d = 0.25 * (r23 + r32) / c
            a = 0.25 * (r13 - r31) / c
        else:
            d = 0.5 * np.sqrt(zd)
            b = 0.25 * (r13 + r31) / d
            c = 0.25 * (r23 + r32) / d
            a = 0.25 * (r21 - r12) / d
        if a < 0.0:
            b, c, d = -b, -c, -d
    return float(b), float(c), float(d), float(qx), float(qy), float(qz), float(dx), float(dy), float(dz), qfac


def _open(path: str, mode: str):
    return gzip.open(path, mode) if path.endswith(".gz") else open(path, mode)


def read_nifti(path: str) -> Tuple[np.ndarray, Dict]:
    """Returns (data in file order [x, y, z, ...] with slope / intercept applied as float64 like `get_fdata()`, header dict
    with `dim`, `pixdim`, `qform_code`, `sform_code`, `qform` (4x4 or None), `sform` (4x4 or None), `affine` (best affine),
    `datatype`, `xyzt_units`, `byteorder`)."""
    with _open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 348:
        raise ValueError(f"{path}: shorter than a NIfTI-1 header")
    bo = "<"
    if struct.unpack("<i", raw[:4])[0] != 348:
        bo = ">"
        if struct.unpack(">i", raw[:4])[0] != 348:
            raise ValueError(f"{path}: sizeof_hdr is not 348 in either byte order (not a NIfTI-1 file)")
    magic = raw[344:348]
    if magic[:3] not in (b"n+1", b"ni1"):
        raise ValueError(f"{path}: bad NIfTI-1 magic {magic!r}")
    if magic[:3] == b"ni1":
        raise ValueError(f"{path}: header / image pairs (.hdr + .img) are not supported, only single-file .nii")
    dim = struct.unpack(bo + "8h", raw[40:56])
    datatype, bitpix = struct.unpack(bo + "hh", raw[70:74])
    pixdim = np.array(struct.unpack(bo + "8f", raw[76:108]), np.float64)
    vox_offset, slope, inter = struct.unpack(bo + "3f", raw[108:120])
    xyzt_units = raw[123]
    qform_code, sform_code = struct.unpack(bo + "hh", raw[252:256])
    qb, qc, qd, qx, qy, qz = struct.unpack(bo + "6f", raw[256:280])
    srow = np.array(struct.unpack(bo + "12f", raw[280:328]), np.float64).reshape(3, 4)
    if not 1 <= dim[0] <= 7:
        raise ValueError(f"{path}: dim[0] = {dim[0]}")
    if datatype not in _DTYPES:
        raise ValueError(f"{path}: unsupported NIfTI datatype code {datatype}")
    shape = tuple(int(d) for d in dim[1 : 1 + dim[0]])
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(bo)
    n = int(np.prod(shape))
    off = int(vox_offset) if vox_offset >= 352 else 352
    data = np.frombuffer(raw, dtype=dt, count=n, offset=off).reshape(shape, order="F")
    data = data.astype(np.float64)
    if slope != 0 and np.isfinite(slope) and np.isfinite(inter) and not (slope == 1 and inter == 0):
        data = data * slope + inter
    qfac = -1.0 if pixdim[0] < 0 else 1.0
    qform = quatern_to_mat44(qb, qc, qd, qx, qy, qz, pixdim[1], pixdim[2], pixdim[3], qfac) if qform_code > 0 else None
    sform = np.vstack([srow, [0, 0, 0, 1]]) if sform_code > 0 else None
    if sform is not None:
        affine = sform
    elif qform is not None:
        affine = qform
    else:  # neither: spacings on the diagonal, origin at the centre voxel
        zooms = np.array([abs(p) if p != 0 else 1.0 for p in pixdim[1:4]])
        affine = np.diag(list(zooms) + [1.0])
        full = np.array(list(shape[:3]) + [1] * (3 - len(shape[:3])), np.float64)
        affine[:3, 3] = -(full - 1) / 2 * zooms
    hdr = dict(dim=np.array(dim), pixdim=pixdim, datatype=int(datatype), bitpix=int(bitpix), qform_code=int(qform_code),
               sform_code=int(sform_code), qform=qform, sform=sform, affine=affine, xyzt_units=int(xyzt_units), byteorder=bo,
               scl_slope=float(slope), scl_inter=float(inter))
    return data, hdr


def write_nifti(path: str, data: np.ndarray, affine: Optional[np.ndarray] = None) -> None:
    """Writes `data` ([x, y, z], any real dtype -> stored as float32 unless it already is one of the supported integer
    types) with `affine` as both qform (code 2) and sform (code 1), millimetre units."""
    data = np.asarray(data)
    if data.ndim != 3:
        raise ValueError("write_nifti: expects a 3-D array")
    if data.dtype.name not in _CODES or data.dtype == np.float64:
        data = data.astype(np.float32)
    affine = np.eye(4) if affine is None else np.asarray(affine, np.float64)
    b, c, d, qx, qy, qz, dx, dy, dz, qfac = mat44_to_quatern(affine)
    hdr = bytearray(352)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, data.shape[0], data.shape[1], data.shape[2], 1, 1, 1, 1)
    struct.pack_into("<hh", hdr, 70, _CODES[data.dtype.name], data.dtype.itemsize * 8)
    struct.pack_into("<8f", hdr, 76, qfac, dx, dy, dz, 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<3f", hdr, 108, 352.0, 1.0, 0.0)  # vox_offset, scl_slope, scl_inter
    hdr[123] = 2  # xyzt_units: NIFTI_UNITS_MM
    struct.pack_into("<hh", hdr, 252, 2, 1)  # qform_code "aligned", sform_code "scanner"
    struct.pack_into("<6f", hdr, 256, b, c, d, qx, qy, qz)
    struct.pack_into("<12f", hdr, 280, *affine[:3, :].reshape(-1))
    hdr[344:348] = b"n+1\x00"
    with _open(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(np.asfortranarray(data).astype(data.dtype.newbyteorder("<")).tobytes(order="F"))
