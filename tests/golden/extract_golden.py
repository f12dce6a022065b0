"""Extract the reference's golden OCP result into a plain .npz fixture.

Source (read-only, only present in the build container):
  /root/reference/agimus_controller/tests/resources/simple_ocp_croco_results.pkl
  (compared by agimus_controller/tests/test_ocp_croco_base.py:175-204)

The pickle is UNTRUSTED content, so it is never un-pickled: the opcode stream is
walked with ``pickletools.genops`` and only the BYTEARRAY8 / BINBYTES payloads
(little-endian f64 buffers of numpy arrays) are kept.  SURVEY.md §8(c): arrays
0-9 = states (14), 10-18 = Riccati gains (98 -> 7x14 C order), 19-27 = feed
forward terms (7).

Run:  python tests/golden/extract_golden.py
Writes tests/golden/simple_ocp_croco_results.npz
"""
import pathlib
import pickletools
import sys

import numpy as np

SRC = pathlib.Path(
    "/root/reference/agimus_controller/tests/resources/simple_ocp_croco_results.pkl"
)
DST = pathlib.Path(__file__).parent / "simple_ocp_croco_results.npz"


def main() -> int:
    raw = SRC.read_bytes()
    payloads = []
    floats = []
    for op, arg, _pos in pickletools.genops(raw):
        if op.name in ("BYTEARRAY8", "BINBYTES", "BINBYTES8", "SHORT_BINBYTES"):
            if len(arg) % 8 == 0 and len(arg) >= 56:
                payloads.append(np.frombuffer(bytes(arg), dtype="<f8").copy())
        elif op.name == "BINFLOAT":
            floats.append(arg)
    if len(payloads) == 28:
        states = np.stack(payloads[0:10])
        gains = np.stack(payloads[10:19]).reshape(9, 7, 14)
        ffs = np.stack(payloads[19:28])
    else:
        # ``.tolist()`` pickles: a flat stream of BINFLOATs, 10*14 + 9*98 + 9*7.
        f = np.asarray(floats, dtype=np.float64)
        assert f.size == 10 * 14 + 9 * 98 + 9 * 7, (len(payloads), f.size)
        states = f[:140].reshape(10, 14)
        gains = f[140 : 140 + 882].reshape(9, 7, 14)
        ffs = f[140 + 882 :].reshape(9, 7)
    assert states.shape == (10, 14) and gains.shape == (9, 7, 14) and ffs.shape == (9, 7)
    # spot values quoted in SURVEY.md §8(c)
    assert abs(states[9][0] - (-0.0467182276)) < 1e-9
    assert abs(ffs[0][0] - (-1965.9911627546)) < 1e-6
    assert abs(gains[0][0, 0] - 686.02444604) < 1e-6
    np.savez(DST, states=states, ricatti_gains=gains, feed_forward_terms=ffs)
    print("wrote", DST, states.shape, gains.shape, ffs.shape)
    return 0


if __name__ == "__main__":
    sys.exit(main())
