"""Plane store: the reference's `.par` scene files, read and written with its safe_saving /
safe_loading protocol (nerf_helpers.py:19-67, models.py:612-678), host staging, and the torch.distributed broadcast.
Host logic only — the device staging (`to_device` / `attach`) needs a GPU: tests/test_gpu_next_rows.py."""
import json
import os
import shutil
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
import nvsr_b200  # noqa: F401
from nvsr_b200 import plane_store as PS

REF = os.environ.get("NVSR_REFERENCE", "/root/reference")
SCENE = "tiny_DS2_PlRes6_4"


@pytest.fixture
def store_dir(tmp_path):
    shutil.copy(os.path.join(H.GOLDEN, "coarse_%s.par" % SCENE), tmp_path)
    return str(tmp_path)


def _twin():
    return H.golden("coarse_%s_par_twin.npz" % SCENE)


def test_reads_a_file_the_reference_wrote(store_dir):
    rec = PS.PlaneStore(store_dir).read(SCENE)
    twin = _twin()
    assert sorted(rec.planes) == PS.plane_names(SCENE) == sorted(k for k in twin if k != "box")
    for k, v in rec.planes.items():
        assert v.dtype == torch.float32 and np.array_equal(v.numpy(), twin[k])
    assert np.array_equal(torch.as_tensor(rec.box).numpy(), twin["box"])
    assert rec.opt_states == [None] * 4 and rec.nbytes() == 4 * (3 * 8 * 36 + 8 * 16)


def test_write_protocol_and_fallbacks(store_dir):
    st = PS.PlaneStore([os.path.join(store_dir, "missing"), store_dir])     # list of locations: first hit wins
    rec = st.read(SCENE)
    other = {k: v + 1.0 for k, v in rec.planes.items()}
    f = st.write(SCENE, other, rec.box)                                       # overwrite through temp/bckp
    assert f == os.path.join(store_dir, "coarse_%s.par" % SCENE)
    assert sorted(os.listdir(store_dir)) == ["coarse_%s.par" % SCENE]         # neither _temp nor _bckp left behind
    assert torch.equal(st.read(SCENE).planes[PS.plane_names(SCENE)[0]], other[PS.plane_names(SCENE)[0]])
    st.write(SCENE, rec.planes, rec.box, as_best=True)
    assert os.path.isfile(f + "_best") and st.path(SCENE, prefer_best=True) == f
    assert torch.equal(st.read(SCENE, prefer_best=True).planes[PS.plane_names(SCENE)[1]], rec.planes[PS.plane_names(SCENE)[1]])
    # a torn main file: the loader falls back to _temp, then to _bckp (nerf_helpers.py:53-66)
    shutil.copy(f, f + "_bckp")
    with open(f, "wb") as fh:
        fh.write(b"torn")
    assert torch.equal(st.read(SCENE).planes[PS.plane_names(SCENE)[2]], other[PS.plane_names(SCENE)[2]])
    shutil.copy(f + "_best", f + "_temp")
    assert torch.equal(st.read(SCENE).planes[PS.plane_names(SCENE)[2]], rec.planes[PS.plane_names(SCENE)[2]])
    os.remove(f + "_temp"), os.remove(f + "_bckp")
    with pytest.raises(Exception):
        st.read(SCENE)
    with pytest.raises(FileNotFoundError):
        st.read("no_such_scene")


def test_prefetch_and_host_cache(store_dir):
    st = PS.PlaneStore(store_dir)
    st.prefetch(SCENE)
    st.prefetch(SCENE)                       # idempotent while pending
    rec = st.host_record(SCENE)
    assert st.host_record(SCENE) is rec      # cached
    st.prefetch("no_such_scene")
    with pytest.raises(FileNotFoundError):   # the background error surfaces on use
        st.host_record("no_such_scene")
    st.evict(SCENE)
    assert st.host_record(SCENE) is not rec
    with pytest.raises(RuntimeError):
        st.to_device(SCENE, "cpu")           # no CPU render path


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present on this machine")
def test_reference_loads_what_we_write(store_dir):
    """our safe_saving -> the reference's safe_loading (nerf_helpers.py:50-67), plain and _best"""
    st = PS.PlaneStore(store_dir)
    rec = st.read(SCENE)
    planes = {k: v * 2.0 - 0.25 for k, v in rec.planes.items()}
    f = st.write("other_DS2_PlRes6_4", planes, rec.box, opt_states=[None, None, None, None])
    st.write("other_DS2_PlRes6_4", planes, rec.box, as_best=True)
    for extra in ([], ["best"]):
        res = subprocess.run([sys.executable, os.path.join(H.GOLDEN, "make_golden_planestore.py"), "--load", f] + extra,
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr[-3000:]
        s = json.loads([l for l in res.stdout.splitlines() if l.startswith("PLANESTORE_JSON ")][-1][len("PLANESTORE_JSON "):])
        assert s["keys"] == ["coords_normalization", "opt_states", "params"] and s["n_opt_states"] == 4
        assert np.array_equal(np.array(s["box"]), torch.as_tensor(rec.box).double().numpy())
        for k, v in planes.items():
            assert s["planes"][k][0] == list(v.shape)
            assert s["planes"][k][1] == float(v.double().sum()) and s["planes"][k][2] == float(v.double().abs().max())


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _bcast_worker(rank, world, port, store_dir, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # only rank 1 can see the files; rank 0 receives everything over the process group
        st = PS.PlaneStore(store_dir if rank == 1 else os.path.join(store_dir, "nothing_here"))
        planes, box = st.broadcast(SCENE, src=1)
        assert st.host_record(SCENE).planes is planes      # the received record is cached: no file access afterwards
        torch.save({"planes": planes, "box": box}, os.path.join(out_dir, "rank%d.pt" % rank))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_broadcast(store_dir, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("bcast"))
    mp.spawn(_bcast_worker, args=(2, _free_port(), store_dir, out), nprocs=2, join=True)
    a, b = torch.load(os.path.join(out, "rank0.pt")), torch.load(os.path.join(out, "rank1.pt"))
    twin = _twin()
    for k in PS.plane_names(SCENE):
        assert torch.equal(a["planes"][k], b["planes"][k]) and np.array_equal(a["planes"][k].numpy(), twin[k])
    assert torch.equal(a["box"], b["box"]) and np.array_equal(a["box"].numpy(), twin["box"])
