"""CPU tests of the data formats either side of the path: the numpy NIfTI-1 reader / writer
(nesvor_b200/image/nifti.py, restated from the NIfTI-1 standard because nibabel is absent), the affine <-> slice-transform
geometry against golden vectors produced by the REFERENCE's own image_utils.py (tests/golden/make_golden_affine.py), and
the save / load functions of image.py round-tripping stacks, volumes and slice folders."""
import gzip
import os
import struct

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "affine_ref.npz")


def _rot(rng):
    q, r = np.linalg.qr(rng.normal(size=(3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 2] = -q[:, 2]
    return q


def test_affine_geometry_matches_reference_goldens():
    from nesvor_b200.image import affine2transformation, compare_resolution_affine, transformation2affine
    from nesvor_b200.transform import RigidTransform

    g = np.load(GOLD)
    for i in range(int(g["n_cases"])):
        vol = torch.tensor(g[f"a2t_{i}_vol"])
        v2, m2, tr = affine2transformation(vol, vol > 0, g[f"a2t_{i}_res"], g[f"a2t_{i}_affine"])
        assert np.array_equal(v2.numpy(), g[f"a2t_{i}_vol_out"]) and np.array_equal(m2.numpy(), g[f"a2t_{i}_mask_out"])
        np.testing.assert_allclose(tr.matrix(True).numpy(), g[f"a2t_{i}_mat"], rtol=1e-5, atol=1e-4)
        aff = transformation2affine(vol, RigidTransform(torch.tensor(g[f"t2a_{i}_mat"]), True), *[float(r) for r in g[f"a2t_{i}_res"]])
        np.testing.assert_allclose(aff, g[f"t2a_{i}_affine"], rtol=1e-5, atol=1e-4)
    r, a = np.array([1.0, 1.0, 3.0]), np.eye(4)
    got = [compare_resolution_affine(r, a, r + 5e-4, a, (3, 4, 5), (3, 4, 5)), compare_resolution_affine(r, a, r + 2e-3, a, (3, 4, 5), (3, 4, 5)),
           compare_resolution_affine(r, a, r, a + 2e-3, (3, 4, 5), (3, 4, 5)), compare_resolution_affine(r, a, r, a, (3, 4, 5), (3, 4, 6))]
    assert got == [bool(x) for x in g["compare"]] == [True, False, False, False]


def test_affine_round_trip_and_voxel_positions():
    """affine -> per-slice transforms -> world position of every voxel == affine @ (i, j, k, 1), right- and left-handed."""
    from nesvor_b200.image import affine2transformation

    rng = np.random.default_rng(0)
    for left in (False, True):
        d, h, w = 4, 5, 6
        res = np.array([0.8, 1.2, 3.0])
        M = _rot(rng) @ np.diag(res)
        if left:
            M[:, 0] = -M[:, 0]
        A = np.eye(4)
        A[:3, :3], A[:3, 3] = M, rng.uniform(-30, 30, 3)
        vol = torch.arange(d * h * w, dtype=torch.float32).view(d, h, w)
        v2, _, tr = affine2transformation(vol, vol > -1, res, A)
        mat = tr.matrix(True).numpy().astype(np.float64)
        for k in range(d):
            for (j, i) in ((0, 0), (h - 1, w - 1), (2, 3)):
                i_file = (w - 1 - i) if left else i  # the image was mirrored along x
                assert float(v2[k, j, i]) == float(vol[k, j, i_file])
                x_slice = np.array([(i - (w - 1) / 2) * res[0], (j - (h - 1) / 2) * res[1], 0.0])
                world = mat[k, :, :3] @ (x_slice + mat[k, :, 3])
                np.testing.assert_allclose(world, (A @ np.array([i_file, j, k, 1.0]))[:3], atol=1e-4)
        assert np.linalg.det(mat[0, :, :3]) > 0.999


def test_nifti_header_layout_and_round_trip(tmp_path):
    from nesvor_b200.image.nifti import mat44_to_quatern, quatern_to_mat44, read_nifti, write_nifti

    rng = np.random.default_rng(1)
    data = rng.normal(size=(5, 6, 7)).astype(np.float32)
    A = np.eye(4)
    A[:3, :3] = _rot(rng) @ np.diag([0.7, 0.9, 2.5])
    A[:3, 3] = [-12.5, 30.25, 4.0]
    for name in ("a.nii", "a.nii.gz"):
        p = str(tmp_path / name)
        write_nifti(p, data, A)
        raw = gzip.open(p, "rb").read() if name.endswith(".gz") else open(p, "rb").read()
        # the published NIfTI-1 layout (nifti1.h), field by field
        assert struct.unpack("<i", raw[0:4])[0] == 348 and raw[344:348] == b"n+1\x00"
        assert struct.unpack("<8h", raw[40:56]) == (3, 5, 6, 7, 1, 1, 1, 1)
        assert struct.unpack("<hh", raw[70:74]) == (16, 32)  # DT_FLOAT32, bitpix
        pixdim = struct.unpack("<8f", raw[76:108])
        assert pixdim[0] == 1.0 and np.allclose(pixdim[1:4], [0.7, 0.9, 2.5], atol=1e-6)
        assert struct.unpack("<f", raw[108:112])[0] == 352.0 and raw[123] == 2  # vox_offset, mm
        assert struct.unpack("<hh", raw[252:256]) == (2, 1)  # qform "aligned", sform "scanner"
        assert np.allclose(np.array(struct.unpack("<12f", raw[280:328])).reshape(3, 4), A[:3], atol=1e-5)
        assert len(raw) == 352 + data.size * 4
        assert np.array_equal(np.frombuffer(raw, "<f4", data.size, 352).reshape(data.shape, order="F"), data)  # x fastest
        back, hdr = read_nifti(p)
        assert np.array_equal(back.astype(np.float32), data)
        np.testing.assert_allclose(hdr["affine"], A, atol=1e-5)
        np.testing.assert_allclose(hdr["qform"], A, atol=1e-5)  # the quaternion path reproduces the same affine
        assert hdr["qform_code"] == 2 and hdr["sform_code"] == 1 and hdr["xyzt_units"] == 2
    # quaternion <-> matrix on its own, including qfac = -1 and the 180-degree branch
    for left in (False, True):
        for R in (_rot(rng), np.diag([1.0, -1.0, -1.0]), np.diag([-1.0, -1.0, 1.0])):
            M = np.eye(4)
            M[:3, :3] = R @ np.diag([1.5, 0.5, 2.0])
            if left:
                M[:3, 2] = -M[:3, 2]
            M[:3, 3] = [1, 2, 3]
            b, c, d, qx, qy, qz, dx, dy, dz, qfac = mat44_to_quatern(M)
            assert qfac == (-1.0 if left else 1.0)
            np.testing.assert_allclose(quatern_to_mat44(b, c, d, qx, qy, qz, dx, dy, dz, qfac), M, atol=1e-9)


def test_nifti_reader_dtypes_scaling_byte_order_and_affine_choice(tmp_path):
    from nesvor_b200.image.nifti import read_nifti

    def make(path, arr, bo="<", code=4, slope=2.0, inter=-1.0, qcode=0, scode=0, quat=(0, 0, 0, 0, 0, 0), srow=None, pixdim=(1, 2, 3, 4), ext=0):
        hdr = bytearray(352 + ext)
        struct.pack_into(bo + "i", hdr, 0, 348)
        struct.pack_into(bo + "8h", hdr, 40, 3, *arr.shape, 1, 1, 1, 1)
        struct.pack_into(bo + "hh", hdr, 70, code, arr.dtype.itemsize * 8)
        struct.pack_into(bo + "8f", hdr, 76, *pixdim, 1, 1, 1, 1)
        struct.pack_into(bo + "3f", hdr, 108, 352.0 + ext, slope, inter)
        struct.pack_into(bo + "hh", hdr, 252, qcode, scode)
        struct.pack_into(bo + "6f", hdr, 256, *quat)
        if srow is not None:
            struct.pack_into(bo + "12f", hdr, 280, *np.asarray(srow, np.float32).reshape(-1))
        hdr[344:348] = b"n+1\x00"
        with open(path, "wb") as f:
            f.write(bytes(hdr) + arr.astype(arr.dtype.newbyteorder(bo)).tobytes(order="F"))

    arr = (np.arange(24, dtype=np.int16).reshape(2, 3, 4) - 5)
    for bo in ("<", ">"):
        p = str(tmp_path / f"i16{bo == '<'}.nii")
        make(p, arr, bo=bo, ext=16)  # int16, slope / intercept, a 16-byte header extension before the voxels
        data, hdr = read_nifti(p)
        assert np.array_equal(data, arr * 2.0 - 1.0) and hdr["byteorder"] == bo
        # neither sform nor qform: spacings on the diagonal, origin at the centre voxel
        np.testing.assert_allclose(hdr["affine"], [[2, 0, 0, -1.0], [0, 3, 0, -3.0], [0, 0, 4, -6.0], [0, 0, 0, 1]])
    p = str(tmp_path / "u8.nii")
    make(p, np.arange(24, dtype=np.uint8).reshape(2, 3, 4), code=2, slope=0.0, inter=7.0)  # slope 0 = no scaling
    assert np.array_equal(read_nifti(p)[0], np.arange(24).reshape(2, 3, 4))
    # qform only (90 degrees about z: b = c = 0, d = sin 45), then sform wins when both are present
    s = np.sqrt(0.5)
    p = str(tmp_path / "q.nii")
    make(p, arr, qcode=1, quat=(0, 0, s, 10, 20, 30), pixdim=(1, 2, 3, 4))
    np.testing.assert_allclose(read_nifti(p)[1]["affine"], [[0, -3, 0, 10], [2, 0, 0, 20], [0, 0, 4, 30], [0, 0, 0, 1]], atol=1e-6)
    srow = [[1, 0, 0, 5], [0, 1, 0, 6], [0, 0, 1, 7]]
    make(p, arr, qcode=1, scode=2, quat=(0, 0, s, 10, 20, 30), srow=srow)
    np.testing.assert_allclose(read_nifti(p)[1]["affine"][:3], srow)
    make(p, arr, qcode=1, quat=(0, 0, s, 0, 0, 0), pixdim=(-1, 2, 3, 4))  # qfac = -1 flips the z column
    np.testing.assert_allclose(read_nifti(p)[1]["affine"][:3, 2], [0, 0, -4], atol=1e-6)
    with open(p, "wb") as f:
        f.write(b"\x00" * 400)
    with pytest.raises(ValueError):
        read_nifti(p)


def test_save_load_stack_volume_and_slices(tmp_path):
    from nesvor_b200.image import Slice, Volume, load_slices, load_stack, load_volume, save_nii_volume, save_slices
    from nesvor_b200.transform import RigidTransform

    rng = np.random.default_rng(2)
    d, h, w = 5, 6, 7
    res = (0.9, 1.1, 3.0)
    img = torch.tensor(rng.uniform(0.1, 1.0, size=(d, h, w)), dtype=torch.float32)
    mat = torch.tensor(np.concatenate([_rot(rng), rng.uniform(-20, 20, (3, 1))], -1)[None], dtype=torch.float32)
    vol = Volume(img, img > 0.5, RigidTransform(mat, True), *res)
    pv, pm = str(tmp_path / "vol.nii.gz"), str(tmp_path / "mask.nii.gz")
    vol.save(pv, masked=False)
    Volume(vol.mask.float(), None, vol.transformation, *res).save(pm, masked=False)
    # volume: image, mask, spacings and the volume-centred transform survive
    back = load_volume(pv, pm)
    assert torch.equal(back.image, img) and torch.equal(back.mask, vol.mask)
    assert np.allclose([back.resolution_x, back.resolution_y, back.resolution_z], res, atol=1e-6)
    np.testing.assert_allclose(back.transformation.matrix(True).numpy(), mat.numpy(), atol=2e-4)
    # stack: one transform per z-slice, slice k's centre = volume transform applied to (0, 0, (k - (d-1)/2) s_z)
    st = load_stack(pv, pm)
    assert len(st) == d and st.slices.shape == (d, 1, h, w) and st.thickness == pytest.approx(3.0) and st.gap == pytest.approx(3.0)
    sm = st.transformation.matrix(True).numpy()
    R, t = mat[0, :, :3].numpy(), mat[0, :, 3].numpy()
    for k in range(d):
        centre_world = R @ (np.array([0, 0, (k - (d - 1) / 2) * res[2]]) + t)
        np.testing.assert_allclose(sm[k, :, :3] @ sm[k, :, 3], centre_world, atol=2e-4)
    sl = st[2]
    assert isinstance(sl, Slice) and sl.image.shape == (1, h, w) and len(st[1:3]) == 2
    # a mask with a different grid is refused
    other = Volume(vol.mask.float(), None, vol.transformation, 0.9, 1.1, 2.0)
    other.save(pm, masked=False)
    with pytest.raises(Exception, match="do not match"):
        load_stack(pv, pm)
    # slice folder: ids order the slices, masked pixels are written as zeros, poses survive
    slices = [st[k] for k in (3, 0, 4)]
    folder = str(tmp_path / "slices")
    save_slices(folder, slices)
    assert sorted(os.listdir(folder)) == ["0.nii.gz", "1.nii.gz", "2.nii.gz"]
    loaded = load_slices(folder)
    for a, b in zip(slices, loaded):
        assert torch.equal(b.image, a.image * a.mask) and torch.equal(b.mask, (a.image * a.mask) > 0)
        np.testing.assert_allclose(b.transformation.matrix(True).numpy(), a.transformation.matrix(True).numpy(), atol=2e-4)
        assert b.resolution_z == pytest.approx(3.0)
    # 4-D input with a singleton channel and a 4-D file are handled like the reference does
    save_nii_volume(str(tmp_path / "c.nii"), img[:, None], None)
    assert load_volume(str(tmp_path / "c.nii")).image.shape == (d, h, w)


def test_nifti_quaternion_convention_against_scipy():
    """NIfTI's (a, b, c, d) is the unit quaternion (w, x, y, z) with a >= 0 recovered from b, c, d (nifti1.h: "quatern_b,
    quatern_c, quatern_d"); checked against scipy's independent implementation, with spacings and qfac applied per column."""
    from scipy.spatial.transform import Rotation

    from nesvor_b200.image.nifti import mat44_to_quatern, quatern_to_mat44

    rng = np.random.default_rng(5)
    for _ in range(20):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        if q[0] < 0:
            q = -q
        a, b, c, d = q
        dx, dy, dz = rng.uniform(0.5, 3.0, 3)
        for qfac in (1.0, -1.0):
            M = quatern_to_mat44(b, c, d, 1.0, 2.0, 3.0, dx, dy, dz, qfac)
            R = Rotation.from_quat([b, c, d, a]).as_matrix()  # scipy: scalar-last
            np.testing.assert_allclose(M[:3, :3], R @ np.diag([dx, dy, dz * qfac]), atol=1e-12)
            np.testing.assert_allclose(M[:3, 3], [1.0, 2.0, 3.0])
            b2, c2, d2, *_rest, qf2 = mat44_to_quatern(M)
            assert qf2 == qfac
            np.testing.assert_allclose([b2, c2, d2], [b, c, d], atol=1e-9)
