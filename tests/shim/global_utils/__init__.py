"""TEST-ONLY stand-in for the reference author's un-vendored `global_utils` package (github.com/caprilovel/
global_utils, not pinned, not installable offline).  It provides exactly the names the
reference's drivers import (main.py:1-3,14-16; denoise_train.py:9,91; Transfer_learning.py:1,14-16) with the
semantics inferred from those call sites, so that `main.py` and `denoise_train.py` can run UNCHANGED in the drop-in
tests.  Not part of the product."""
