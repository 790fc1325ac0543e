"""g16b200 -- host-side mirror of codex-storage/nim-groth16's prover interface over libg16b200.so,
the sm_100a CUDA backend.  Module map (reference module -> here):
    groth16/prover.nim       -> prover.py      groth16/bn128/msm.nim   -> bn128.py
    groth16/math/ntt.nim     -> ntt.py         groth16/zkey_types.nim  -> zkey_types.py
    groth16/files/*.nim      -> files.py       groth16/fake_setup.nim  -> fake_setup.py
    groth16/files/export_json.nim -> export_json.py
parallel.py adds the multi-GPU split.  Nothing here computes on the CPU."""
from . import _lib  # noqa: F401
from .encoding import FORM_MONT, FORM_STD  # noqa: F401
from .zkey_types import JENS_GROTH, SNARKJS, Mask, Proof, R1CS, Witness, ZKey  # noqa: F401
from .bn128 import (fixed_base_g1, fixed_base_g2, msm_g1, msm_g2, msm_multi_threaded_g1,  # noqa: F401
                    msm_multi_threaded_g2)
from .ntt import Domain, create_domain, forward_ntt, inverse_ntt  # noqa: F401
from .prover import (ProverContext, build_abc, compute_quotient_pointwise, compute_snarkjs_scalar_coeffs,  # noqa: F401
                     generate_proof, generate_proof_with_mask, generate_proof_with_trivial_mask)
from .fake_setup import (ToxicWaste, create_fake_circuit_setup, fake_circuit_setup, r1cs_to_coeffs,  # noqa: F401
                         random_toxic_waste, synthetic_chain_circuit)
from . import files, export_json, parallel  # noqa: F401
