"""ORACLE package — test infrastructure only (see ref_forward.py).  Never imported by pytorchcv_b200/."""
from .ref_forward import oracle_forward, OracleUnsupported  # noqa: F401
from .seeded import seeded_init, seeded_input  # noqa: F401
from .bf16_storage import oracle_forward_bf16_storage, oracle_forward_16bit_storage  # noqa: F401
