#!/bin/bash
# usage: tools/probes/build_variant.sh <tag> [-DFLAG=V ...]   ->  tools/probes/lib_<tag>.so (developer experiments)
TAG=$1; shift
python - "$TAG" "$@" <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from footsies_gym_b200 import build
tag, flags = sys.argv[1], sys.argv[2:]
print(build.build(force=True, extra_flags=flags, lib_path=os.path.join(os.getcwd(), "tools", "probes", f"lib_{tag}.so")))
PY
