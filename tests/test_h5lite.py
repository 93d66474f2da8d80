"""himo_b200.h5lite -- the self-contained HDF5-subset codec behind the `.h5` scene files.

Pins:
  * READER against a file written by the REAL HDF5 library: tests/golden/libhdf5_written_testhdf5_7.4.mat is scipy's
    test asset `testhdf5_7.4_GLNX86.mat` (a MATLAB v7.3 file, i.e. HDF5 behind a 512-byte user block, written by libhdf5
    in 2008: superblock v0, symbol-table group, version-1 object header, layout v2, contiguous float64).  scipy's own
    test-suite documents its content: `testdouble` = 0 : pi/4 : 2 pi.
  * WRITER through that reader; its dataset headers carry the same message types (and, for the fill-value message and the
    float64 datatype, the same bytes) as the library-written file; datasets inserted IN PLACE into a copy of the
    library-written file leave its own content intact.
  * the `.h5`-backed frame store end to end: dataset assembly, run_save writing `<res_name>` into the scene files (the
    reference's `r+` delete-and-create, OSF/src/trainer.py:337-343), save_zip and eval reading it back -- equal to the
    `.npy`-per-array store on the same data.
"""
import os
import shutil
import struct

import numpy as np
import pytest

from conftest import GOLDEN
from himo_b200 import h5lite, himo, runner, store
from himo_b200.dataset import HDF5Dataset

LIBFILE = os.path.join(GOLDEN, "libhdf5_written_testhdf5_7.4.mat")


def test_reads_a_file_written_by_the_real_library():
    with h5lite.File(LIBFILE) as f:
        assert f.base == 512 and f.leaf_k == 4 and f.internal_k == 16          # user block, default B-tree ranks
        assert f.keys() == ["testdouble"] and "testdouble" in f and "nope" not in f
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.float64
        np.testing.assert_allclose(d[...].ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)   # the known answer
        with pytest.raises(KeyError):
            f["missing"]


def test_written_headers_carry_the_library_messages(tmp_path):
    p = str(tmp_path / "w.h5")
    with h5lite.File(p, "w") as f:
        f.create_group("g").create_dataset("x", data=np.arange(9, dtype=np.float64).reshape(9, 1))
    with h5lite.File(p) as mine, h5lite.File(LIBFILE) as lib:
        a = dict(mine["g/x"]._o.msgs)
        b = dict(lib["testdouble"]._o.msgs)
        assert {0x0001, 0x0003, 0x0005, 0x0008} <= set(a) and {0x0001, 0x0003, 0x0005, 0x0008} <= set(b)
        assert a[0x0003][:20] == b[0x0003][:20]          # IEEE little-endian float64 datatype message, byte for byte
        assert a[0x0005][:8] == b[0x0005][:8]            # fill-value message, byte for byte
        assert a[0x0001][:24] == b[0x0001][:24]          # dataspace v1, rank 2, dims (9, 1)
        assert a[0x0008][0] == 3 and b[0x0008][0] == 2   # layout: v3 (what libhdf5 >= 1.8 / h5py writes) vs the 2008 file's v2
    raw = open(p, "rb").read()
    assert raw[:8] == h5lite.SIGNATURE and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)                     # end-of-file address


def test_round_trip_every_dtype_of_the_schema_and_many_members(tmp_path):
    rng = np.random.default_rng(0)
    p = str(tmp_path / "scene.h5")
    arrays = {"lidar": rng.normal(size=(100, 4)).astype(np.float32), "ground_mask": rng.random(100) > 0.5,
              "pose": np.eye(4), "pose32": np.eye(4, dtype=np.float32), "lidar_id": rng.integers(0, 5, 100).astype(np.uint8),
              "flow_instance_id": rng.integers(-3, 900, 100).astype(np.int16), "scania_instance": rng.integers(0, 2 ** 31, 7).astype(np.uint32),
              "lidar_dt": rng.random(100).astype(np.float32), "empty": np.zeros((0, 3), np.float32), "i64": np.arange(5),
              "half": np.arange(4).astype(np.float16), "scalar": np.float32(3.5)}
    stamps = [str(315969904359876000 + 100000 * k) for k in range(160)]           # more groups than one B-tree leaf holds
    with h5lite.File(p, "w") as f:
        for ts in stamps:
            g = f.create_group(ts)
            for name, arr in arrays.items():
                g.create_dataset(name, data=arr)
    with h5lite.File(p) as f:
        assert f.keys() == sorted(stamps) and len(f.keys()) == 160
        for ts in (stamps[0], stamps[77], stamps[-1]):
            assert f[ts].keys() == sorted(arrays)
            for name, arr in arrays.items():
                got = f[ts][name][...]
                assert got.dtype == np.asarray(arr).dtype and got.shape == np.asarray(arr).shape and (got == arr).all(), name


def test_in_place_insert_replace_delete_also_in_a_library_written_file(tmp_path):
    q = str(tmp_path / "lib.h5")
    shutil.copy(LIBFILE, q)
    with h5lite.File(q, "r+") as f:                                              # a file h5lite did not write
        f.create_dataset("flow", data=np.arange(12, dtype=np.float32).reshape(4, 3))
        f.create_group("315969904359876000").create_dataset("ground_mask", data=np.array([True, False, True]))
        for j in range(14):                                                      # enough names to grow the heap and split a node
            f.create_dataset(f"label_{j:02d}", data=np.arange(j + 1, dtype=np.int16))
        f.create_dataset("flow", data=np.full((4, 3), 7, np.float32))            # replace = the reference's del + create
        del f["label_03"]
    with h5lite.File(q) as f:
        assert f.keys() == sorted(f.keys()) and "label_03" not in f and "label_13" in f
        np.testing.assert_allclose(f["testdouble"][...].ravel(), np.arange(9) * np.pi / 4, atol=1e-15)   # untouched
        assert (f["flow"][...] == 7).all() and f["315969904359876000/ground_mask"][...].tolist() == [True, False, True]
        assert (f["label_09"][...] == np.arange(10)).all()
    with pytest.raises(h5lite.H5Error):
        with h5lite.File(q, "r") as f:
            f.create_dataset("nope", data=np.zeros(3))


def test_unsupported_files_fail_loudly(tmp_path):
    p = str(tmp_path / "v2.h5")
    open(p, "wb").write(h5lite.SIGNATURE + bytes([2]) + bytes(100))
    with pytest.raises(h5lite.H5Unsupported):
        h5lite.File(p)
    open(p, "wb").write(b"not hdf5" * 100)
    with pytest.raises(h5lite.H5Error):
        h5lite.File(p)
    with h5lite.File(str(tmp_path / "w.h5"), "w") as f:
        with pytest.raises(h5lite.H5Unsupported):
            f.create_dataset("s", data=np.array(["a", "b"]))


class _ConstEngine:
    """Stands in for the GPU engine: final flow = 0.01 * index pattern (the store path is what is under test)."""

    def infer(self, item):
        n = item["pc0"].shape[0]
        return (np.arange(n * 3, dtype=np.float32).reshape(n, 3) % 7) * 0.01

    def infer_stream(self, items):
        for it in items:
            yield self.infer(it)


def test_h5_store_pipeline_equals_npy_store(tmp_path):
    d_npy, d_h5 = str(tmp_path / "av2_npy"), str(tmp_path / "av2_h5")
    os.makedirs(d_h5)
    store.write_synthetic_dataset(d_npy, n_scenes=2, n_frames=6, n_points=1500, seed=3)
    h5 = store.H5Store(d_h5, backend=h5lite)
    store.write_synthetic_dataset(d_h5, n_scenes=2, n_frames=6, n_points=1500, seed=3, store=h5)
    assert sorted(f for f in os.listdir(d_h5) if f.endswith(".h5")) and isinstance(store.open_store(d_h5), store.H5Store)
    a, b = HDF5Dataset(d_npy, n_frames=3), HDF5Dataset(d_h5, n_frames=3)
    assert isinstance(b.store, store.H5Store) and len(a) == len(b)
    for i in (0, 3, len(a) - 1):
        x, y = a[i], b[i]
        assert x.keys() == y.keys()
        for k in x:
            if isinstance(x[k], np.ndarray):
                assert x[k].dtype == y[k].dtype and np.array_equal(x[k], y[k]), k
    for d in (d_npy, d_h5):
        assert runner.run_save({"dataset_path": d, "res_name": "fake_flow"}, engine=_ConstEngine(), n_frames=3) > 0
        runner.run_save({"dataset_path": d, "res_name": "fake_flow"}, engine=_ConstEngine(), n_frames=3)   # re-run: replace
    ra, rb = HDF5Dataset(d_npy, vis_name="fake_flow", eval=True), HDF5Dataset(d_h5, vis_name="fake_flow", eval=True)
    for i in range(len(ra)):
        assert np.array_equal(ra[i]["fake_flow"], rb[i]["fake_flow"])
    za = runner.run_save_zip({"data_dir": d_npy, "res_name": "fake_flow"})
    zb = runner.run_save_zip({"data_dir": d_h5, "res_name": "fake_flow"})
    key = (ra[0]["scene_id"], str(ra[0]["timestamp"]))
    assert np.array_equal(himo.read_output_zip(za, key), himo.read_output_zip(zb, key))
    ea = runner.run_eval({"data_dir": d_npy, "res_name": "fake_flow", "out_json": str(tmp_path / "a.json"), "metrics_device": "host"})
    eb = runner.run_eval({"data_dir": d_h5, "res_name": "fake_flow", "out_json": str(tmp_path / "b.json"), "metrics_device": "host"})
    assert ea == eb
