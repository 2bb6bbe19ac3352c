#!/usr/bin/env python3
"""`python src/meta_overlay.py` / `make meta_overlay`: the reference's entry point (src/meta_overlay.py),
served by the B200-native implementation in ecseg_b200/."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecseg_b200.meta_overlay import main  # noqa: E402

if __name__ == "__main__":
    main(sys.argv[1:])
