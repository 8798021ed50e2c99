"""The product path under torch.distributed: two ranks (gloo rendezvous, both on cuda:0) run lec_fixed on
their time shards; the gathered results equal the single-process run bit for bit and only rank 0 writes."""
import argparse
import logging
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import helpers as H

pytestmark = pytest.mark.gpu
INP = os.path.join(H.GOLDEN, "inputs")
SAM = os.path.join(H.GOLDEN, "samples")


def _args():
    return argparse.Namespace(infile=os.path.join(SAM, "Catarina_NCEP-R2.nc"), fixed=True, track=False, choose=False,
                              residuals=True, box_limits=None, outname=None, plots=False, cdsapi=False, mpas=False)


def _run(outdir, box_file):
    from lorenzcycletoolkit_b200.frameworks import lec_fixed
    from lorenzcycletoolkit_b200.utils import preprocessing as PP
    a = _args()
    a.box_limits = box_file
    nl = PP.read_namelist(os.path.join(INP, "namelist_NCEP-R2"))
    data = PP.prepare_data(a, os.path.join(INP, "namelist_NCEP-R2"), box_limits_file=box_file)
    os.makedirs(os.path.join(outdir, "lv"), exist_ok=True)
    return lec_fixed(data, nl, outdir, os.path.join(outdir, "lv"), logging.getLogger("t"), a,
                     engine_options={"device": 0})


def _worker(rank, world, port, outdir, box_file, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK="0")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    df = _run(os.path.join(outdir, f"rank{rank}"), box_file)
    q.put((rank, df.values.copy()))
    dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path):
    box_file = str(tmp_path / "box")
    with open(box_file, "w") as f:
        f.write("min_lon;-55\nmax_lon;-36\nmin_lat;-35\nmax_lat;-20\n")
    single = _run(str(tmp_path / "single"), box_file)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), box_file, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert np.array_equal(got[r], single.values), r
    assert os.path.exists(tmp_path / "rank0" / "Catarina_NCEP-R2_fixed_results.csv")
    assert not os.path.exists(tmp_path / "rank1" / "Catarina_NCEP-R2_fixed_results.csv")
    assert len(os.listdir(tmp_path / "rank1" / "lv")) == 0 and len(os.listdir(tmp_path / "rank0" / "lv")) == 21


def _cli_run(workdir, argv, nproc, env_extra=None):
    """Run the root CLI module in ``workdir`` -- plainly (nproc = 1) or under torchrun (one rank per GPU when the
    box has nproc GPUs: NCCL; several ranks on one GPU otherwise: gloo, chosen by ``cli.init_distributed``)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    env.update(env_extra or {})
    script = os.path.join(root, "lorenzcycletoolkit.py")
    if nproc == 1:
        cmd = [sys.executable, script] + argv
    else:
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), script] + argv
    r = subprocess.run(cmd, cwd=workdir, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r


def _tree(d):
    out = {}
    for base, _, files in os.walk(d):
        for f in files:
            if f.startswith("log."):
                continue
            p = os.path.join(base, f)
            out[os.path.relpath(p, d)] = open(p, "rb").read()
    return out


@pytest.mark.parametrize("mode,nproc", [("fixed", 2), ("track", 2), ("fixed", 4)])
def test_cli_under_torchrun_writes_one_identical_set_of_files(tmp_path, mode, nproc):
    """`torchrun --nproc-per-node N lorenzcycletoolkit.py X.nc -r -f|-t`: cli.main joins the process group itself,
    the ranks shard the time steps (LEC terms and the 850-hPa diagnostics), rank 0 alone writes, and every file is
    byte-identical to the single-process run.  Four ranks on the five time steps of the file leave the last rank
    with an EMPTY shard (ceil(5 / 4) = 2 steps per rank)."""
    import shutil
    import torch
    for name in ("one", "two"):
        d = tmp_path / name / "inputs"
        os.makedirs(d)
        shutil.copy(os.path.join(INP, "namelist_NCEP-R2"), d / "namelist")
        shutil.copy(os.path.join(INP, "box_limits_Reg1"), d / "box_limits")
        shutil.copy(os.path.join(INP, "track_testdata_NCEP-R2"), d / "track")
    argv = [os.path.join(SAM, "testdata_NCEP-R2.nc"), "-r", "-f" if mode == "fixed" else "-t"]
    _cli_run(str(tmp_path / "one"), argv, 1)
    r = _cli_run(str(tmp_path / "two"), argv, nproc)
    one, two = _tree(tmp_path / "one" / "LEC_Results"), _tree(tmp_path / "two" / "LEC_Results")
    assert sorted(one) == sorted(two) and len(one) >= 22
    for k in one:
        assert one[k] == two[k], k
    # the log file has one writer too (rank 0): a single "Starting" line
    logs = [p for p in (tmp_path / "two" / "LEC_Results").rglob("log.*")]
    assert len(logs) == 1 and logs[0].read_text().count("Starting LEC analysis") == 1
    if torch.cuda.device_count() >= nproc:
        # one rank per GPU: the NCCL branch of the result gather ran (cli.init_distributed picks it)
        _cli_run(str(tmp_path / "two"), argv, nproc, {"LEC_DIST_BACKEND": "nccl"})
        again = _tree(tmp_path / "two" / "LEC_Results")
        for k in one:
            assert one[k] == again[k], k


def test_reused_handle_grows_its_staging_window():
    """A handle whose first lec_run_host saw one or two slots must not keep that window: a later multi-step call
    on the same handle needs slot_m, slot, slot_p resident (LEC_ERR_NOMEM although memory is free)."""
    from oracle import lec_oracle as O
    P, _ = H.load_prepared("Catarina_NCEP-R2.nc")
    box = (-55, -36, -35, -20)
    P = O.slice_domain_fixed(P, *box)
    fields, scale = H.engine_inputs(P, np.float64)
    steps = H.fixed_steps(P, *box)
    with H.make_engine(P, np.float64, scale, max_steps=64) as eng:
        full, _, _ = eng.run_host(fields, steps)
    with H.make_engine(P, np.float64, scale, max_steps=64) as eng:
        one = steps[:1].copy()
        one["slot"], one["slot_m"], one["slot_p"] = 0, 0, 0           # a single resident slot
        one["ct_m"], one["ct_0"], one["ct_p"] = 0.0, 0.0, 0.0
        eng.run_host([f[:1] for f in fields], one)
        again, _, _ = eng.run_host(fields, steps)                       # needs a 3+-slot window
    assert np.array_equal(full, again)
