"""snarkjs-compatible proof.json / public.json writers (groth16/files/export_json.nim:25-80)."""
from __future__ import annotations

from .encoding import fr_from_std, g1_from_array, g2_from_array
from .zkey_types import Proof


def public_io_json(prf: Proof) -> str:
    """exportPublicIO (export_json.nim:25-45): skips the leading constant 1."""
    vals = fr_from_std(prf.publicIO)
    assert len(vals) > 0 and vals[0] == 1                          # export_json.nim:30-31
    lines = []
    for i, v in enumerate(vals[1:], start=1):
        lines.append(("[ " if i == 1 else ", ") + '"%d"' % v)
    lines.append("] ")
    return "\n".join(lines) + "\n"


def proof_json(prf: Proof) -> str:
    """exportProof (export_json.nim:70-80): decimal strings, projective z = 1."""
    (ax, ay), = g1_from_array(prf.pi_a)
    (bx, by), = g2_from_array(prf.pi_b)
    (cx, cy), = g1_from_array(prf.pi_c)
    g1 = lambda x, y: '    [ "%d"\n    , "%d"\n    , "1"\n    ]\n' % (x, y)
    fp2 = lambda c, z: '    %s [ "%d"\n      , "%d"\n      ]\n' % (c, z[0], z[1])
    return ('{ "protocol": "groth16"\n, "curve":    "bn128"\n, "pi_a":\n' + g1(ax, ay) + ', "pi_b":\n'
            + fp2("[", bx) + fp2(",", by) + fp2(",", (1, 0)) + "    ]\n" + ', "pi_c":\n' + g1(cx, cy) + "}\n")


def export_proof(fpath: str, prf: Proof) -> None:
    with open(fpath, "w") as f:
        f.write(proof_json(prf))


def export_public_io(fpath: str, prf: Proof) -> None:
    with open(fpath, "w") as f:
        f.write(public_io_json(prf))
