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
