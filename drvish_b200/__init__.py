"""drvish_b200: B200-native (sm_100a) implementation of drvish's per-minibatch ELBO step."""
from .elbo import FusedTraceELBO
from .models import DRNBVAE, MCVNBVAE, NBVAE
from .optim import FusedAggMo, PyroFusedAggMo, PyroFusedScheduler
from .modules import Encoder, LinearMultiBias, NBDecoder, make_fc, split_in_half

__all__ = ["NBVAE", "DRNBVAE", "MCVNBVAE", "FusedTraceELBO", "FusedAggMo", "PyroFusedAggMo", "PyroFusedScheduler", "Encoder", "NBDecoder", "LinearMultiBias", "make_fc", "split_in_half"]
