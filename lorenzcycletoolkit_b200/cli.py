"""Command line of the B200 engine, argument-compatible with the reference's
``lorenzcycletoolkit.py`` (:50-129 parser, :132-155 results layout, :158-200 dispatch):

    python -m lorenzcycletoolkit_b200.cli samples/testdata_NCEP-R2.nc -r -f
    python -m lorenzcycletoolkit_b200.cli samples/testdata_NCEP-R2.nc -r -t --trackfile inputs/track

Inputs (``inputs/namelist``, ``inputs/box_limits``, track files) and outputs
(``./LEC_Results/<stem>_<method>/...``) are those of the reference.  ``--choose`` and ``--cdsapi``
(interactive GUI, network download) are outside the engine's scope and are rejected with a clear
message; ``--plots`` is accepted (the reference's own tests pass it) and answered with a warning: the
figures are the reference's ``src/plots`` applied to the CSVs written here."""

from __future__ import annotations

import argparse
import logging
import os
import sys
import time

from .frameworks import lec_fixed, lec_moving
from .utils.preprocessing import prepare_data, read_namelist


def create_arg_parser():
    p = argparse.ArgumentParser(description="Lorenz Energy Cycle (LEC) program -- B200 engine.")
    p.add_argument("infile", help="Input .nc file with temperature, geopotential/geopotential height, "
                                  "and wind components data.")
    p.add_argument("-r", "--residuals", action="store_true",
                   help="Compute the Dissipation and Generation terms as residuals.")
    g = p.add_mutually_exclusive_group(required=True)
    g.add_argument("-f", "--fixed", action="store_true",
                   help="Compute the energetics for a fixed domain specified by the 'box_limits' file.")
    g.add_argument("-t", "--track", action="store_true", help="Define the domain using a track file.")
    g.add_argument("-c", "--choose", action="store_true", help="Interactively select the domain for each time step.")
    p.add_argument("-z", "--zeta", action="store_true",
                   help="Use the vorticity from the track file instead of computing it at 850 hPa.")
    p.add_argument("-m", "--mpas", action="store_true",
                   help="Specify this flag if working with MPAS-A data processed with MPAS-BR routines.")
    p.add_argument("-p", "--plots", action="store_true", help="Generate plots.")
    p.add_argument("-v", "--verbosity", action="store_true", help="Logger level set to debug mode.")
    p.add_argument("--cdsapi", action="store_true", help="Use CDS API for downloading data (experimental).")
    p.add_argument("--time-resolution", type=int, default=3)
    p.add_argument("--trackfile", type=str, default="inputs/track",
                   help="Specify a custom track file. Default is 'inputs/track'.")
    p.add_argument("--box_limits", type=str, default="inputs/box_limits",
                   help="Specify a custom box limits file. Default is 'inputs/box_limits'.")
    p.add_argument("-o", "--outname", type=str, help="Specify an output name for the results.")
    p.add_argument("--namelist", type=str, default="inputs/namelist",
                   help="Variable namelist (the reference hard-codes inputs/namelist).")
    return p


def setup_results_directory(args, method):
    """``./LEC_Results/<infile stem>_<method>/{results_vertical_levels,Figures}``."""
    sub = os.path.join("./LEC_Results/", "".join(args.infile.split("/")[-1].split(".nc")) + "_" + method)
    levels = os.path.join(sub, "results_vertical_levels")
    figures = os.path.join(sub, "Figures")
    for d in (figures, sub, levels):
        os.makedirs(d, exist_ok=True)
    return sub, figures, levels


def initialize_logging(results_subdirectory, args):
    """Logger ``lorenzcycletoolkit`` (DEBUG with -v, else INFO) to ``log.<stem>`` and the console
    (tools.py:32-73)."""
    verbose = bool(args.verbosity)
    logging.basicConfig(level=logging.INFO if verbose else logging.ERROR,
                        format="%(asctime)s - %(levelname)s - %(message)s")
    log = logging.getLogger("lorenzcycletoolkit")
    level = logging.DEBUG if verbose else logging.INFO
    log.setLevel(level)
    log.propagate = False
    for h in list(log.handlers):
        log.removeHandler(h)
    fmt = logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s")
    from .sharding import is_rank0
    ch = logging.StreamHandler()
    handlers = [ch]
    if is_rank0():
        # under torchrun only rank 0 owns log.<stem> (mode "w" would truncate it once per rank)
        handlers.append(logging.FileHandler(
            os.path.join(results_subdirectory, f'log.{os.path.basename(args.infile).split(".")[0]}'), mode="w"))
    else:
        level = max(level, logging.WARNING)          # the other ranks only report problems, on the console
        log.setLevel(level)
    for h in handlers:
        h.setLevel(level)
        h.setFormatter(fmt)
        log.addHandler(h)
    return log


def run_lec_analysis(data, args, results_subdirectory, figures_directory,
                     results_subdirectory_vertical_levels, app_logger, namelist="inputs/namelist"):
    start = time.time()
    variable_list_df = read_namelist(namelist)
    if getattr(args, "plots", False):
        app_logger.warning("⚠️  --plots: figures are produced by the reference's src/plots from the CSVs written "
                           "here; this engine does not draw them.")
    if args.fixed:
        df = lec_fixed(data, variable_list_df, results_subdirectory, results_subdirectory_vertical_levels,
                       app_logger, args)
        app_logger.info("🎉 Analysis complete! Fixed framework ran in %.2f seconds" % (time.time() - start))
        return df
    df = lec_moving(data, variable_list_df, None, results_subdirectory, figures_directory,
                    results_subdirectory_vertical_levels, app_logger, args)
    app_logger.info("🎉 Analysis complete! Moving framework ran in %.2f seconds" % (time.time() - start))
    return df


def init_distributed():
    """Join the ``torch.distributed`` job when launched by ``torchrun`` (RANK / WORLD_SIZE / LOCAL_RANK in
    the environment): one process per GPU, time steps sharded over the ranks, one all-gather of the
    per-step results, rank 0 writes the files.  NCCL when every rank has its own GPU,
    gloo otherwise (several ranks sharing one GPU).  Returns True when this call created the group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return False
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return False
    local = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
    ngpu = torch.cuda.device_count()
    if ngpu == 0:
        raise RuntimeError("torchrun launch without a CUDA device: the B200 engine has no CPU path")
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    backend = os.environ.get("LEC_DIST_BACKEND") or ("nccl" if ngpu >= local_world else "gloo")
    torch.cuda.set_device(local % ngpu)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local % ngpu))
    else:
        dist.init_process_group(backend)
    return True


def main(argv=None):
    args = create_arg_parser().parse_args(argv)
    if args.choose:
        sys.exit("--choose needs an interactive matplotlib display and is outside the B200 engine's scope")
    if args.cdsapi:
        sys.exit("--cdsapi downloads data over the network and is outside the B200 engine's scope")
    created = init_distributed()
    try:
        method = "fixed" if args.fixed else "track"
        sub, figures, levels = setup_results_directory(args, method)
        log = initialize_logging(sub, args)
        log.info("⏳ Starting LEC analysis (B200 engine)")
        data = prepare_data(args, args.namelist, log,
                            box_limits_file="inputs/box_limits" if os.path.exists("inputs/box_limits") else args.box_limits)
        return run_lec_analysis(data, args, sub, figures, levels, log, namelist=args.namelist)
    finally:
        if created:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
