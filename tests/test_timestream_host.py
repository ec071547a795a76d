"""SURVEY section 8 (f4), host side: the m-mode transform of a timestream, the noise realisation of
simulate() and the Timestream bookkeeping against a fixture produced by the reference's own
drift/pipeline/timestream.py (tests/golden/make_golden_timestream.py).  No GPU needed."""

import os
import types

import numpy as np
import pytest

SMALL_CFG = dict(
    num_freq=3, freq_start=100.0, freq_end=112.0, freq_mode="edge",
    num_cylinders=2, cylinder_width=5.0, num_feeds=3, feed_spacing=1.5, tsys=1.0,
)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "timestream_small.npz"))


def _manager(tmp_path):
    from driftscan_b200.core import beamtransfer
    from driftscan_b200.telescope import cylinder

    tel = cylinder.PolarisedCylinderTelescope.from_config(SMALL_CFG)
    return types.SimpleNamespace(beamtransfer=beamtransfer.BeamTransfer(str(tmp_path / "bt"), telescope=tel))


def _write_timestream(ts, data):
    from driftscan_b200.util import h5lite

    for fi in range(data.shape[0]):
        os.makedirs(ts._fdir(fi), exist_ok=True)
        with h5lite.File(ts._ffile(fi), "w") as f:
            f.create_dataset("timestream", data=data[fi])
            f.attrs["ntime"] = data.shape[-1]


def test_generate_mmodes_against_reference(tmp_path, gold):
    from driftscan_b200.pipeline import timestream

    m = _manager(tmp_path)
    ts = timestream.Timestream(str(tmp_path / "ts"), m)
    _write_timestream(ts, gold["timestream"])
    assert ts.ntime == int(gold["ntime"])
    assert np.array_equal(ts.timestream_f(1), gold["timestream"][1])
    ts.generate_mmodes()
    tel = m.beamtransfer.telescope
    assert tel.mmax == int(gold["mmax"])
    for mi in range(tel.mmax + 1):
        mm = ts.mmode(mi)
        assert mm.shape == (tel.nfreq, 2, tel.npairs)
        assert np.abs(mm - gold["mmodes"][mi]).max() <= 1e-13 * np.abs(gold["mmodes"]).max()
    assert (ts.mmode(0)[:, 1] == 0).all()  # the negative half of m = 0 stays empty
    assert os.path.exists(str(tmp_path / "ts" / "mmodes" / "COMPLETED_M"))
    ts.generate_mmodes()  # second call: nothing to do


def test_simulate_noise_against_reference(tmp_path, gold):
    from driftscan_b200.pipeline import timestream

    m = _manager(tmp_path)
    ts = timestream.simulate(m, str(tmp_path / "noise"), maps=[], ndays=int(gold["ndays"]), resolution=900.0,
                             seed=int(gold["seed"]) + 7)
    tel = m.beamtransfer.telescope
    assert ts.ntime == int(gold["noise_ntime"]) == 96
    got = np.array([ts.timestream_f(fi) for fi in range(tel.nfreq)])
    assert np.abs(got - gold["noise_timestream"]).max() <= 1e-13 * np.abs(gold["noise_timestream"]).max()
    from driftscan_b200.util import h5lite

    with h5lite.File(ts._ffile(2), "r") as f:
        assert np.allclose(np.array(f["phi"][:]), np.linspace(0, 2 * np.pi, 96, endpoint=False))
        assert np.array_equal(np.array(f["feedmap"][:]), tel.feedmap)
        assert np.array_equal(np.array(f["uniquepairs"][:]), tel.uniquepairs)
        assert np.array_equal(np.array(f["baselines"][:]), tel.baselines)
        assert f.attrs["beamtransfer_path"] == os.path.abspath(m.beamtransfer.directory)
    # the object comes back from its pickle
    ts2 = timestream.Timestream.load(str(tmp_path / "noise"))
    assert ts2.directory == ts.directory and ts2.output_directory == ts.output_directory


WORKER = r"""
import os, sys, types
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from test_timestream_host import SMALL_CFG, _write_timestream
from driftscan_b200 import parallel
from driftscan_b200.core import beamtransfer
from driftscan_b200.pipeline import timestream
from driftscan_b200.telescope import cylinder

dist.init_process_group("gloo")
comm = parallel.Comm.current()
base = sys.argv[2]
gold = np.load(os.path.join(sys.argv[1], "tests", "golden", "timestream_small.npz"))
tel = cylinder.PolarisedCylinderTelescope.from_config(SMALL_CFG)
m = types.SimpleNamespace(beamtransfer=beamtransfer.BeamTransfer(os.path.join(base, "bt"), telescope=tel))
ts = timestream.Timestream(os.path.join(base, "ts"), m)
if comm.rank0:
    _write_timestream(ts, gold["timestream"])
comm.barrier()
ts.generate_mmodes()   # frequencies split for the FFT, m split for the files
ok = True
for mi in range(tel.mmax + 1):
    ok &= bool(np.abs(ts.mmode(mi) - gold["mmodes"][mi]).max() <= 1e-13 * np.abs(gold["mmodes"]).max())
# noise-only simulation: every rank writes its own frequencies, seeded with seed + rank
tsn = timestream.simulate(m, os.path.join(base, "noise"), maps=[], ndays=5, seed=11)
comm.barrier()
shapes = [tsn.timestream_f(fi).shape for fi in range(tel.nfreq)]
ok &= all(s == (tel.npairs, 2 * tel.mmax + 1) for s in shapes)
lo, hi = comm.split_range(tel.nfreq)
np.random.seed(11 + comm.rank)
nz = (np.array([1.0, 1.0j]) * np.random.standard_normal((tel.npairs, hi - lo, 2 * tel.mmax + 1, 2))).sum(axis=-1)
nps = tel.noisepower(np.arange(tel.npairs)[:, None], np.arange(lo, hi)[None, :], ndays=5).reshape(tel.npairs, hi - lo)
want = np.fft.ifft(nz * (nps[:, :, None] / 2.0) ** 0.5, axis=-1) * (2 * tel.mmax + 1)
for lfi, fi in enumerate(range(lo, hi)):
    ok &= bool(np.allclose(tsn.timestream_f(fi), want[:, lfi], rtol=1e-12, atol=0))
comm.barrier()
print("RANK", comm.rank, "OK" if ok else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


@pytest.mark.timeout(300)
def test_two_ranks_gloo(tmp_path):
    import socket
    import subprocess
    import sys

    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2", OMP_NUM_THREADS="1")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script), root, str(tmp_path)], env=e, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"RANK {r} OK" in o, o
