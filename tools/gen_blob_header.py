#!/usr/bin/env python3
"""Regenerate the enum blocks of include/cosim_blob.h from cosim_b200/model.py (DIMS / OPTS)."""
import os, re, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from cosim_b200.model import DIMS, OPTS

path = os.path.join(os.path.dirname(__file__), "..", "include", "cosim_blob.h")
src = open(path).read()
dim = "enum cosim_dim {\n" + "".join(f"  CD_{d} = {i},\n" for i, d in enumerate(DIMS)) + "  CD__count\n};"
opt = "enum cosim_opt {\n" + "".join(f"  CO_{d} = {i},\n" for i, d in enumerate(OPTS)) + "  CO__count\n};"
src = re.sub(r"enum cosim_dim \{.*?\};", dim, src, flags=re.S)
src = re.sub(r"enum cosim_opt \{.*?\};", opt, src, flags=re.S)
open(path, "w").write(src)
print("updated", path)
