"""Oracle (test infrastructure): numpy restatement of replay transition assembly.

Follows ``/root/reference/rainbowiqn/redis_memory.py``:
  * get_byte_multiple_transition (which 7 ring slots a sample reads)   redis_memory.py:347-369
  * get_torch_tensor_from_byte_transition (blank masking, n-step return, stacking)  :479-541

The Redis hash per slot (timestep, state bytes, action, reward, nonterminal; :184-193) becomes
five parallel arrays indexed by the global slot id = actor*actor_capacity + position.
"""
import numpy as np


class ReplayStore:
    def __init__(self, actor_capacity, nb_actor):
        self.actor_capacity = actor_capacity
        cap = actor_capacity * nb_actor
        self.timestep = np.zeros(cap, np.int64)
        self.frame = np.zeros((cap, 84, 84), np.uint8)
        self.action = np.zeros(cap, np.int64)
        self.reward = np.zeros(cap, np.float64)
        self.nonterminal = np.zeros(cap, np.bool_)

    def write(self, id_actor, actor_index, timesteps, frames, actions, rewards, dones):
        """Frame half of append_actor_buffer                           redis_memory.py:174-199"""
        n = len(actions)
        pos = (np.arange(actor_index, actor_index + n) % self.actor_capacity) + id_actor * self.actor_capacity
        self.timestep[pos] = timesteps
        self.frame[pos] = frames
        self.action[pos] = actions
        self.reward[pos] = rewards
        self.nonterminal[pos] = ~np.asarray(dones, np.bool_)
        return pos

    def window_slots(self, data_index, history=4, n_step=3):
        """Slots idx-history+1 .. idx+n within the actor's ring       redis_memory.py:349-366"""
        cap = self.actor_capacity
        actor = data_index // cap
        return [((k + data_index - history + 1) % cap) + actor * cap for k in range(history + n_step)]

    def assemble(self, data_indexes, history=4, n_step=3, discount=0.99):
        """Returns states u8 (B,4,84,84), actions i64, returns f32, next_states u8, nonterminals f32.

        The reference then converts frames to fp32 and divides by 255 (:527-536); the uint8 stacks are
        returned here so both the oracle (``/255``) and the CUDA path (uint8 ingest) can consume them.
        """
        states, nexts, acts, rets, nts = [], [], [], [], []
        for d in data_indexes:
            slots = self.window_slots(int(d), history, n_step)
            ts = [int(self.timestep[s]) for s in slots]
            nt = [bool(self.nonterminal[s]) for s in slots]
            rw = [float(self.reward[s]) for s in slots]
            fr = [self.frame[s] for s in slots]
            blank = np.zeros((84, 84), np.uint8)
            # blank_trans = Transition(0, zeros, None, 0, False)              :12
            for t in range(history - 2, -1, -1):                             # :494-496
                if ts[t + 1] == 0:
                    ts[t], fr[t], rw[t], nt[t] = 0, blank, 0.0, False
            for t in range(history, history + n_step):                       # :497-499
                if not nt[t - 1]:
                    ts[t], fr[t], rw[t], nt[t] = 0, blank, 0.0, False
            states.append(np.stack(fr[:history]))
            nexts.append(np.stack(fr[n_step:n_step + history]))
            rets.append(sum(discount ** k * rw[history + k - 1] for k in range(n_step)))  # :516-518
            acts.append(int(self.action[slots[history - 1]]))
            nts.append(nt[history + n_step - 1])
        return (np.stack(states), np.array(acts, np.int64), np.array(rets, np.float64).astype(np.float32),
                np.stack(nexts), np.array(nts, np.float32))
