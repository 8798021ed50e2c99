#!/usr/bin/env python
"""Same command line as the reference's ``lorenzcycletoolkit.py`` (e.g.
``python lorenzcycletoolkit.py samples/testdata_NCEP-R2.nc -r -f``), served by the B200 engine."""
from lorenzcycletoolkit_b200.cli import (create_arg_parser, initialize_logging, main,  # noqa: F401
                                         run_lec_analysis, setup_results_directory)
from lorenzcycletoolkit_b200.utils.preprocessing import prepare_data  # noqa: F401

if __name__ == "__main__":
    main()
