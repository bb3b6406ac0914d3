"""ds2i_b200 — B200-native query path over ds2i-format inverted indexes."""
from .api import (Index, WandData, QueryBatch, Group, query_batch, flatten_queries, read_queries, and_query, and_freq_query, or_query,  # noqa: F401
                  or_freq_query, ranked_and_query, wand_query, maxscore_query, ranked_or_query, OPS, RANKED)
