"""PPO.memory: the flat transition store `learn()` consumes (reference: /root/reference/PPO/Memory.py:7-30).

Same surface as the reference - `.states/.actions/.rewards/.dones` behave like Python lists of float32 items, `push`,
`clear` - but transitions that arrive from the device rollout buffer stay in HBM as flat env-major tensors
(states [N][O], actions [N][AW], rewards [N], dones [N]); host items are only materialised if somebody indexes them.
"""
from __future__ import annotations

import numpy as np
import torch

FIELDS = ("states", "actions", "rewards", "dones")


class _Field:
    """List-like view of one field: host items (pushed / `+=`-ed by user code) followed by device rows."""

    def __init__(self, owner, name):
        self._owner, self._name, self._host = owner, name, []

    def __len__(self):
        return len(self._host) + (0 if self._name in self._owner._dev_cleared else self._owner._dev_count)

    def append(self, item):
        self._owner._host_after_device()
        self._host.append(item)

    def extend(self, items):
        for it in items:
            self.append(it)

    def __iadd__(self, items):
        self.extend(items)
        return self

    def _rows(self):
        out = list(self._host)
        t = self._owner._dev.get(self._name)
        if t is not None and self._owner._dev_count and self._name not in self._owner._dev_cleared:
            arr = t[: self._owner._dev_count].cpu().numpy()
            if arr.ndim == 2 and arr.shape[1] == 1 and self._name == "actions" and self._owner._scalar_actions:
                arr = arr[:, 0]
            out.extend(list(arr))
        return out

    def __iter__(self):
        return iter(self._rows())

    def __getitem__(self, i):
        """Items like the reference's list: a single index downloads that one row, not the whole field."""
        nh, nd = len(self._host), len(self) - len(self._host)
        if isinstance(i, (int, np.integer)):
            j = int(i) + (nh + nd if i < 0 else 0)
            if not 0 <= j < nh + nd:
                raise IndexError("memory field index out of range")
            if j < nh:
                return self._host[j]
            row = self._owner._dev[self._name][j - nh].cpu().numpy()
            if self._name == "actions" and row.shape == (1,) and self._owner._scalar_actions:
                row = row[0]
            return row
        return self._rows()[i]

    def __delitem__(self, i):
        if i == slice(None, None, None):
            self.clear()
        else:
            raise TypeError("only `del field[:]` is supported")

    def clear(self):
        """`del memory.states[:]` (the reference's own Memory.clear idiom, Memory.py:26-30) empties the device rows of this
        field too; once all four fields have been emptied the device store is reset."""
        self._host.clear()
        own = self._owner
        if own._dev_count:
            own._dev_cleared.add(self._name)
            if len(own._dev_cleared) == len(FIELDS):
                own._dev_count = 0
                own._dev_cleared.clear()

    def __repr__(self):
        return f"<{self._name}: {len(self)} items>"


class Memory:
    def __init__(self):
        self._dev = {}
        self._dev_count = 0
        self._dev_cap = 0
        self._scalar_actions = True
        self._total = None  # device int64 written by the transfer kernel: rows in the store after the last transfer
        self._expect_total = None   # what the host believes that number is (checked by verify_transfers)
        self._overflow = None       # the rollout buffer's overflow flag of the last transfer
        self._dev_cleared = set()   # fields emptied one by one through `del field[:]`
        self._fields = {n: _Field(self, n) for n in FIELDS}
        # per-row by-products of the fused worker (not part of the reference's Memory): log-prob and state value under the acting
        # policy, GAE returns.  Valid for the first _eval_rows rows as long as _eval_tag (what they were computed from: the acting
        # policy's version, gamma, lambda) still describes the PPO that consumes them - PPO.learn() checks and otherwise recomputes.
        self._eval = {}
        self._eval_rows = 0
        self._eval_tag = None
        self._eval_has_returns = False

    # list-like attributes, assignable like the reference's plain lists
    states = property(lambda s: s._fields["states"], lambda s, v: s._assign("states", v))
    actions = property(lambda s: s._fields["actions"], lambda s, v: s._assign("actions", v))
    rewards = property(lambda s: s._fields["rewards"], lambda s, v: s._assign("rewards", v))
    dones = property(lambda s: s._fields["dones"], lambda s, v: s._assign("dones", v))

    def _assign(self, name, value):
        if value is self._fields[name]:
            return  # `mem.states += [...]` re-assigns the same object
        self._fields[name]._host = list(value)

    def _host_after_device(self):
        if self._dev_count and not self._dev_cleared:
            raise RuntimeError("mixing host pushes after a device transfer is not supported; call learn() or clear() first")

    def push(self, state, action, reward, done):
        """Memory.py:14-24: every item is stored as float32."""
        for name, x in zip(FIELDS, (state, action, reward, done)):
            self._fields[name].append(np.asarray(x).astype(np.float32))

    def clear(self):
        for f in self._fields.values():
            f._host.clear()
        self._dev_count = 0
        self._dev_cleared.clear()
        self._eval_rows, self._eval_tag = 0, None

    def evaluated(self, n, tag):
        """(logp, values, returns or None) of the first n rows if the fused worker produced them for exactly these rows under `tag`."""
        if n and self._eval_rows == n == self._dev_count and self._eval_tag == tag and not self._dev_cleared:
            e = self._eval
            return e["logp"][:n], e["values"][:n], (e["returns"][:n] if self._eval_has_returns else None)
        return None

    def verify_transfers(self):
        """Host-synchronising check that the device agrees with the host's bookkeeping: the transfer kernel's row total equals
        the count the host derived from the rollout's step score (a capacity miscount would otherwise silently truncate the
        batch), and no rollout buffer overflowed.  PPO.learn() calls it where it synchronises anyway."""
        if self._expect_total is not None:
            got = int(self._total.item())
            want, self._expect_total = self._expect_total, None
            if got != want:
                raise RuntimeError(f"Memory: the device transferred {got} rows but the host expected {want}")
        if self._overflow is not None:
            ov, self._overflow = int(self._overflow.item()), None
            if ov:
                raise RuntimeError("Memory: a rollout buffer overflowed its time capacity (transitions were dropped)")

    # ---- device side -------------------------------------------------------------------------------------------
    def reserve(self, capacity, obs_dim, act_width, device):
        """Make room for `capacity` rows in HBM, keeping what is already there."""
        if self._dev and self._dev_cap >= capacity and self._dev["states"].shape[1] == obs_dim:
            return
        capacity = max(int(capacity), 1)
        new = {"states": torch.empty(capacity, obs_dim, dtype=torch.float32, device=device),
               "actions": torch.empty(capacity, act_width, dtype=torch.float32, device=device),
               "rewards": torch.empty(capacity, dtype=torch.float32, device=device),
               "dones": torch.empty(capacity, dtype=torch.float32, device=device)}
        if self._dev and self._dev_count:
            for k in FIELDS:
                new[k][: self._dev_count] = self._dev[k][: self._dev_count]
        if self._eval:
            old, self._eval = self._eval, {k: torch.empty(capacity, dtype=torch.float32, device=device) for k in ("logp", "values", "returns")}
            if self._eval_rows:
                for k in old:
                    self._eval[k][: self._eval_rows] = old[k][: self._eval_rows]
        self._dev, self._dev_cap = new, capacity
        if self._total is None:
            self._total = torch.zeros(1, dtype=torch.int64, device=device)

    def _upload_host(self, obs_dim, act_width, device):
        """Move items pushed on the host (Memory.push / list +=) into the device rows, in order."""
        nh = len(self._fields["states"]._host)
        if not nh:
            return
        if self._dev_count:
            raise RuntimeError("host items and device rows cannot be combined")
        host = {k: np.array(self._fields[k]._host, dtype=np.float32) for k in FIELDS}
        a = host["actions"].reshape(nh, -1)
        self._scalar_actions = host["actions"].ndim == 1
        self.reserve(nh, obs_dim, act_width, device)
        self._dev["states"][:nh] = torch.from_numpy(host["states"].reshape(nh, obs_dim)).to(device)
        self._dev["actions"][:nh] = torch.from_numpy(np.ascontiguousarray(a[:, :act_width])).to(device)
        self._dev["rewards"][:nh] = torch.from_numpy(host["rewards"].reshape(nh)).to(device)
        self._dev["dones"][:nh] = torch.from_numpy(host["dones"].reshape(nh)).to(device)
        for f in self._fields.values():
            f._host.clear()
        self._dev_count = nh
        self._eval_rows, self._eval_tag = 0, None

    def device_view(self, obs_dim, act_width, device):
        """(states [N][O], actions [N][AW], rewards [N], dones [N]) on the device, host items uploaded first."""
        self._upload_host(obs_dim, act_width, device)
        n = self._dev_count
        return tuple(self._dev[k][:n] for k in FIELDS)

    def append_from_rollout(self, buf, n_new, scalar_actions=True, eval_tag=None, with_returns=False):
        """utils.buffer_to_target_buffer_transfer (utils.py:45-50) on the device: the per-env episodes of the rollout
        buffer are concatenated env-major, time-minor behind the rows already stored; the buffer is cleared.
        eval_tag is not None: buf.logp / buf.values (and buf.returns when with_returns) hold the fused worker's by-products for these
        transitions; they travel along and stay usable while every stored row has them under the same tag."""
        device = buf.lengths.device
        self._upload_host(buf.O, buf.AW, device)
        base = self._dev_count
        # room for a full rollout (every env running to the time limit) when that is at most 4 GiB, so that the store
        # is allocated once instead of growing with the episode lengths
        full = buf.E * buf.T
        want = base + (full if full * (buf.O + buf.AW + 2) * 4 <= (4 << 30) else int(n_new))
        self.reserve(max(want, base + int(n_new)), buf.O, buf.AW, device)
        self._scalar_actions = scalar_actions
        extra = ()
        keep = eval_tag is not None and (base == 0 or (self._eval_rows == base and self._eval_tag == eval_tag and self._eval_has_returns == bool(with_returns)))
        if keep:
            if not self._eval or self._eval["logp"].numel() < self._dev_cap:
                old, self._eval = self._eval, {k: torch.empty(self._dev_cap, dtype=torch.float32, device=device) for k in ("logp", "values", "returns")}
                for k in old:
                    self._eval[k][:base] = old[k][:base]
            extra = [(buf.logp, self._eval["logp"]), (buf.values, self._eval["values"])] + ([(buf.returns, self._eval["returns"])] if with_returns else [])
        buf.transfer(self._dev["states"], self._dev["actions"], self._dev["rewards"], self._dev["dones"], base, self._total, extra=extra)
        self._eval_rows, self._eval_tag, self._eval_has_returns = (base + int(n_new), eval_tag, bool(with_returns)) if keep else (0, None, False)
        self._dev_count = base + int(n_new)
        self._expect_total, self._overflow = self._dev_count, buf.overflow
