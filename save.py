#!/usr/bin/env python
"""save.py -- same command line as the reference's save.py, running on the himo_b200 engine.
See himo_b200/runner.py (main_save) for the accepted arguments and the reference lines it follows."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from himo_b200.runner import main_save  # noqa: E402

if __name__ == "__main__":
    main_save()
