"""Forward error correction on the GPU: mirror of `wavecapsdr.dsp.fec` for the P25 framing path."""
from .bch import BCH_63_16_23, bch_decode, bch_decode_batch  # noqa: F401
from .trellis import TrellisDecoder, trellis_decode, trellis_decode_batch, tsbk_decode_batch  # noqa: F401
