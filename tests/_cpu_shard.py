"""CPU stand-in for one rank's shard engine (same shard_* interface as msspe_b200.Engine), built from the pure
Python oracle.  Test infrastructure: lets the gloo tests exercise msspe_b200.distributed.select_sharded without GPUs."""
import numpy as np
import torch

from oracle import kmer_oracle as ko


class CpuShard:
    def __init__(self, records, W, S, w, k):
        self.segs = ko.get_segment_manager(records, W, S, w, k)
        self.post = []
        self.codes = []
        for d in (0, 1):
            m = {}
            for s in self.segs:
                for wd in s.kmers[d]:
                    m.setdefault(ko.encode(wd), []).append(s.index)
            cs = sorted(m)
            self.codes.append(cs)
            self.post.append([m[c] for c in cs])
        self.ignored = [set(), set()]

    def shard_begin(self, d):
        self.ignored[d] = set()

    def shard_codes(self, d):
        return torch.tensor(self.codes[d], dtype=torch.int64)

    def shard_count(self, d):
        f = [sum(1 for g in p if g not in self.ignored[d]) for p in self.post[d]]
        return torch.tensor(f, dtype=torch.int32), sum(f)

    def shard_n_part(self):
        return (max(s.partition_no for s in self.segs) + 1) if self.segs else 0

    def shard_firstpos(self, d, ids, n_part):
        out = np.full((len(ids), n_part), 0xFFFFFFFF, dtype=np.uint32)
        for t, lid in enumerate(ids):
            if lid == 0xFFFFFFFF:
                continue
            for pos, g in enumerate(self.post[d][int(lid)]):
                if g in self.ignored[d]:
                    continue
                p = self.segs[g].partition_no
                if out[t, p] == 0xFFFFFFFF:
                    out[t, p] = pos
        return out

    def shard_apply(self, d, lid, n_part):
        flags = np.zeros(n_part, dtype=np.uint8)
        if lid != 0xFFFFFFFF:
            for g in self.post[d][int(lid)]:
                self.ignored[d].add(g)
                flags[self.segs[g].partition_no] = 1
        return flags
