"""Oracle (test infrastructure): numpy-float64 restatement of the prioritized-replay sum-tree.

Follows ``/root/reference/rainbowiqn/redis_memory.py`` with the Redis string store replaced
by a plain float64 array ``tree`` (2C-1 nodes, implicit heap, leaf of data index d at d+C-1):

  * update_multiple_value / _propagate_multiple_values          redis_memory.py:139-151, 94-105
  * _retrieve_multiple_values (level-synchronous descent)       redis_memory.py:205-229
  * transform_to_valid_tree_indexes                             redis_memory.py:242-264
  * find_multiple_values / total                                redis_memory.py:267-335
  * sample_byte importance weights                              redis_memory.py:465-475
  * update_priorities (power then update)                       redis_memory.py:557-573
  * append_actor_buffer (index arithmetic + priority write)     redis_memory.py:153-202
  * get_current_capacity                                        redis_memory.py:371-390

Arithmetic is float64 (Redis strings round-trip through ``repr(float)``; the real server's
long-double INCRBYFLOAT cannot be exercised offline -- SURVEY.md §8c).  Known reference
quirks are preserved on purpose: a leaf listed twice in one update ends as p1+p2-old; the root
receives ``np.sum(diffs)`` (numpy pairwise summation) instead of the sequential sum; entries
whose ancestor chain reached the root early turn into index -1 and write to a junk key.
"""
import numpy as np


class SumTree:
    def __init__(self, actor_capacity, nb_actor):
        self.actor_capacity = int(actor_capacity)
        self.nb_actor = int(nb_actor)
        self.full_capacity = self.actor_capacity * self.nb_actor
        # initialize_redis_database                               redis_memory.py:61-92
        self.tree = np.zeros(2 * self.full_capacity - 1, np.float64)
        self.index_actor = np.zeros(self.nb_actor, np.int64)
        self.is_full_actor = np.zeros(self.nb_actor, np.int64)
        self.max_priority = 1.0
        self.memory_full = False

    # ------------------------------------------------------------------ update
    def propagate(self, indexes, diff_values):
        """_propagate_multiple_values                             redis_memory.py:94-105"""
        indexes = np.array(indexes, dtype=np.int64).copy()
        diff_values = np.asarray(diff_values, np.float64)
        while np.max(indexes) > 0:
            for j in range(len(indexes)):
                idx = indexes[j]
                if idx > 0:  # ``!= 0``; negative ones hit the junk key "priorities:-1"
                    self.tree[idx] += diff_values[j]
            indexes = (indexes - 1) // 2
        self.tree[0] += np.sum(diff_values)

    def update_multiple_value(self, indeces, priorities):
        """update_multiple_value                                  redis_memory.py:139-151"""
        indeces = np.asarray(indeces, np.int64)
        old = self.tree[indeces].astype(np.float64)
        self.propagate(indeces, priorities - old)
        if np.float64(max(priorities)) > np.float64(self.max_priority):
            self.max_priority = float(np.float64(max(priorities)))

    def update_priorities(self, idxs, priorities, priority_exponent):
        """ReplayRedisMemory.update_priorities                    redis_memory.py:557-573"""
        priorities = np.power(priorities, priority_exponent)
        self.update_multiple_value(idxs, priorities)

    def append_priorities(self, actor_index, id_actor, priorities):
        """Priority/bookkeeping half of append_actor_buffer       redis_memory.py:153-202"""
        n = len(priorities)
        indexes = (np.arange(actor_index, actor_index + n) % self.actor_capacity) + id_actor * self.actor_capacity
        self.update_multiple_value(indexes + self.full_capacity - 1, priorities)
        self.index_actor[id_actor] = (actor_index + n) % self.actor_capacity
        return indexes

    # ------------------------------------------------------------------ sample
    def total(self):
        return float(self.tree[0])

    def retrieve(self, values):
        """_retrieve_multiple_values                              redis_memory.py:205-229"""
        values = np.array(values, np.float64).copy()
        indexes = np.zeros(len(values), dtype=np.int64)
        n_nodes = 2 * self.full_capacity - 1
        while True:
            lefts, rights = 2 * indexes + 1, 2 * indexes + 2
            if np.min(lefts) >= n_nodes:
                return indexes
            for j in range(len(values)):
                if lefts[j] < n_nodes:  # non-power-of-two guard        :217-220
                    left_sum = self.tree[lefts[j]]
                    if values[j] <= left_sum:
                        indexes[j] = lefts[j]
                    else:
                        indexes[j] = rights[j]
                        values[j] = values[j] - left_sum

    def transform_to_valid(self, tree_indexes, history_length, n_step_length):
        """transform_to_valid_tree_indexes                        redis_memory.py:242-264"""
        data = np.array(tree_indexes, np.int64) - self.full_capacity + 1
        cap = self.actor_capacity
        for j in range(len(data)):
            d = data[j]
            actor = d // cap
            dist = (d % cap) - self.index_actor[actor]
            if 0 <= dist <= history_length:
                data[j] = (d + history_length - dist + 1) % cap + actor * cap
            elif -n_step_length <= dist < 0:
                data[j] = (d - n_step_length - dist - 1) % cap + actor * cap
        return data + self.full_capacity - 1

    def find(self, samples, history_length, n_step_length):
        """find_multiple_values with injected ``samples``         redis_memory.py:267-331

        The reference draws samples[i] = random.uniform(i*seg,(i+1)*seg), seg = total/batch, then
        np.random.shuffle (:276-287); callers inject them here to make runs comparable.
        """
        p_total = self.total()
        tree_indexes = self.retrieve(samples)
        tree_indexes = self.transform_to_valid(tree_indexes, history_length, n_step_length)
        data_indexes = tree_indexes - self.full_capacity + 1
        return self.tree[tree_indexes].copy(), data_indexes, tree_indexes, p_total

    def get_current_capacity(self):
        """get_current_capacity                                   redis_memory.py:371-390"""
        if self.memory_full:
            return self.full_capacity
        capacity, full = 0, True
        for a in range(self.nb_actor):
            if int(self.is_full_actor[a]):
                capacity += self.actor_capacity
            else:
                capacity += int(self.index_actor[a])
                full = False
        self.memory_full = full
        return capacity

    def check(self):
        """check_sumtree_correct: max |left+right-parent|         redis_memory.py:107-136"""
        c = self.full_capacity
        par = self.tree[: c - 1]
        return float(np.max(np.abs(self.tree[1 : 2 * c - 1 : 2] + self.tree[2 : 2 * c - 1 : 2] - par))) if c > 1 else 0.0


def stratified_samples(p_total, batch_size, uniforms, perm=None):
    """samples[i] = a + (b-a)*u_i with a=i*seg, b=(i+1)*seg (CPython random.uniform), then an
    optional injected permutation standing in for np.random.shuffle.   redis_memory.py:276-287"""
    seg = p_total / batch_size
    s = np.array([i * seg + ((i + 1) * seg - i * seg) * float(uniforms[i]) for i in range(batch_size)], np.float64)
    return s if perm is None else s[np.asarray(perm)]


def importance_weights(probs, p_total, capacity, priority_weight):
    """sample_byte: w = (capacity * p/p_total)^-beta / max        redis_memory.py:465-475"""
    probs = np.asarray(probs, np.float64) / p_total
    w = (capacity * probs) ** -priority_weight
    return w / w.max()
