#!/usr/bin/env python
"""Drop-in for HM-16.5_Test_LDP/bin/resi_to_cu_depth_LDP.py: start it by hand in the LDP encoder's working
directory (README.md:64-84), it serves the pred_start.sig / command.dat / resi.yuv / state.dat / cu_depth.dat /
pred_end.sig handshake until killed.  No arithmetic here: it calls libethcnn_b200.so, which reads Thr_info.txt,
model_LDP_2000000_qp22~37.dat.* and model_LDP_200000_qp{22,27,32,37}.dat.* from the cwd.  No CPU fallback."""
from __future__ import annotations

import os
import sys


def _binding():
    try:
        from . import binding
        return binding
    except ImportError:
        here = os.path.dirname(os.path.realpath(__file__))
        sys.path.insert(0, here)
        import binding  # type: ignore
        return binding


def main():
    b = _binding()
    with b.EthCnn(os.environ.get("ETHCNN_MODEL_DIR", "."), None, b.MODE_LDP) as net:
        print("ethcnn: predictor initialized.")
        sys.stdout.flush()
        n = net.ldp_serve(".", int(os.environ.get("ETHCNN_DAEMON_MAX_FRAMES", "0")), int(os.environ.get("ETHCNN_DAEMON_IDLE_MS", "0")))
        print("%d frames predicted." % n)
    return 0


if __name__ == "__main__":
    try:
        sys.exit(main())
    except Exception as e:
        sys.stderr.write("resi_to_cu_depth_LDP: %s\n" % (e,))
        sys.exit(1)
