"""gofeature_b200 -- B200-native brute-force cosine search behind goFeature's API.

`gofeature_b200.api` mirrors the reference's Go interfaces (Cache / Set / Block / Buffer)
over the C ABI of libgofeature_b200.so (include/gofeature_b200.h); `gofeature_b200.sharded`
row-shards a set over the ranks of a torch.distributed job.
"""
from . import api  # noqa: F401
from .api import *  # noqa: F401,F403
