#!/usr/bin/env python3
"""development aid: bench.py against another build of the library (TRN_AB_LIB=<path to a libturner_b200*.so>)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turner_b200 import api

if os.environ.get("TRN_AB_LIB"):
    api.use_library(os.environ["TRN_AB_LIB"])
import bench

sys.exit(bench.main())
