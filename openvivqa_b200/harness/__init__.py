from .config import NS, StubVocab, attention_cfg, m4c_cfg, mcan_cfg, xmt_cfg  # noqa: F401
from .models import (Batch, CoAttentionEncoder, CrossModalTransformer, CrossModalityEncoder, DecoderLayer, Encoder,  # noqa: F401
                     GuidedAttentionEncoder, MCAN, bench_forward, bench_library_math, strict_fp32_library_math)
from .m4c import MMF_M4C  # noqa: F401
from .generative import BeamSearch, Captioner, Decoder  # noqa: F401,E402
