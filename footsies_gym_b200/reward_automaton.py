"""The dense-reward automaton of the step kernel, read from the generated table header the kernel is compiled from
(csrc/frame_tables.h, written by tools/gen_kernel_tables.py): footsies.py:388-405 accumulates +-0.3 in Python float64; only
13 cumulative values are reachable, so the kernel carries an index and takes the exact doubles from these tables.

  CUM_VALUES[i]                 the cumulative reward of an episode so far
  CUM_NEXT[i][code]             index after a frame with guard-drop code (bit 0: P1's bar dropped, bit 1: P2's)
  STEP_REWARD[code]             reward of a non-terminal frame
  TERM_REWARD[i_after][code][p2_dead]   reward of the terminal frame: step reward + ((+1 | -1) - cumulative)
Used on the host only by FootsiesEnv._apply_load_fix (the one step after load_battle_state)."""
import os
import re

_HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "frame_tables.h")


def _define(text, name):
    m = re.search(r"#define " + name + r" (\{.*\})\s*$", text, re.M)
    if not m:
        raise RuntimeError(f"{name} not found in {_HEADER}")
    body = m.group(1).replace("{", "[").replace("}", "]")
    body = re.sub(r"-?0x[0-9a-fA-F.]+p[+-]?\d+", lambda h: repr(float.fromhex(h.group())), body)
    return eval(body, {"__builtins__": {}})      # nested lists of numbers only (generated file of this package)


with open(_HEADER) as _f:
    _text = _f.read()
CUM_VALUES = _define(_text, "FT_CUM_VALUES_INIT")
CUM_NEXT = _define(_text, "FT_CUM_NEXT_INIT")
STEP_REWARD = _define(_text, "FT_STEP_REWARD_INIT")
TERM_REWARD = _define(_text, "FT_TERM_REWARD_INIT")
