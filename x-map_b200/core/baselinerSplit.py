"""BaselinerSplit -- host-side (set operations over users; not on the measured path).
Semantics of xmap/core/baselinerSplit.py:6-146: users present in both domains are split BY USER into
test / keep-both / rest with weights (ratio_split, ratio_both, 1 - both); a test user keeps all
source ratings plus `num_left` randomly chosen target ratings in training, the remaining target
ratings become the test set."""
import random

from ..rdd import LocalRDD, records_of


class BaselinerSplit(object):
    def __init__(self, num_left, ratio_split, ratio_both, seed):
        self.num_left = num_left
        self.ratio_split = ratio_split
        self.ratio_both = ratio_both
        self.seed = seed
        random.seed(seed)                       # the reference seeds the global generator (:31)

    def find_overlap_user(self, sourceRDD, targetRDD):
        tgt = {u for u, _ in records_of(targetRDD)}
        return LocalRDD(dict.fromkeys(u for u, _ in records_of(sourceRDD) if u in tgt))

    def distinguish_data(self, overlap_userRDD_bd, dataRDD):
        ov = set(overlap_userRDD_bd.value)
        recs = records_of(dataRDD)
        return (LocalRDD(r for r in recs if r[0] in ov), LocalRDD(r for r in recs if r[0] not in ov))

    def split_data(self, non_overlap_sourceRDD, overlap_sourceRDD, non_overlap_targetRDD, overlap_targetRDD):
        merged = LocalRDD(records_of(overlap_sourceRDD) + records_of(overlap_targetRDD)).reduceByKey(
            lambda a, b: a + b)
        test_part, both, rest = merged.randomSplit(
            [self.ratio_split, self.ratio_both, 1 - self.ratio_split - self.ratio_both], seed=self.seed)
        train = records_of(non_overlap_sourceRDD) + records_of(non_overlap_targetRDD) + \
            both.collect() + rest.collect()
        test = []
        for uid, lines in test_part.collect():
            source = [x for x in lines if "S:" in x[0]]
            target = [x for x in lines if x not in source]
            remain = random.sample(target, self.num_left)
            test.append((uid, [x for x in target if x not in remain]))
            train.append((uid, source + remain))
        return LocalRDD(train), LocalRDD(test)
