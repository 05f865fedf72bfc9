"""ObsEncoder: host-side mirror of thirdparty/IQN/model.py:111-225 whose compute runs in the CUDA kernels.

Same constructor, same attribute / method names (K, forward, get_qvals, save, load, state_dict ...), same on-disk format
(network_params.pth = plain state_dict with the 14 reference key names, constructor_params.json), same initial weights
for a given seed (the nn.Linear layers are instantiated on the CPU in the reference's order after
torch.manual_seed(seed) and copied into the flat parameter vector).  The parameters live in ONE flat CUDA tensor.
"""
import json
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib, iqn_ops


class ObsEncoder:
    def __init__(self, state_size, action_size, seed, device="cuda:0"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.MarinenavError("ObsEncoder runs on a CUDA device only (there is no CPU fallback); got " + str(device))
        assert state_size == 26, "observation dimension needs to be 26 (velocity, goal, measurements)"   # model.py:124
        assert action_size == iqn_ops.N_ACTIONS
        self.device = dev
        self.seed_id = seed
        self.seed = torch.manual_seed(seed)                       # model.py:117 -- also what makes local == target at start
        self.K = 32
        self.state_size, self.action_size = state_size, action_size
        layers = [nn.Linear(2, 16), nn.Linear(2, 16), nn.Linear(22, 176), nn.Linear(64, 208),
                  nn.Linear(208, 64), nn.Linear(64, 64), nn.Linear(64, action_size)]       # model.py:125-136 order
        flat = torch.cat([t.detach().reshape(-1) for l in layers for t in (l.weight, l.bias)]).float()
        assert flat.numel() == iqn_ops.N_PARAMS
        self.flat = flat.to(dev).contiguous()
        self.packed = torch.empty(iqn_ops.N_PACKED, dtype=torch.float32, device=dev)
        self.packed_tc = torch.empty(iqn_ops.packed_tc_bytes(), dtype=torch.uint8, device=dev)   # bf16 tiles for act_tc
        self.training = True
        self.tau_generator = None                                 # None: CPU default generator, like model.py:149
        self.repack()

    # ---- parameters ------------------------------------------------------------------------------------------------
    def repack(self):
        with torch.cuda.device(self.device):
            iqn_ops.pack(self.flat, self.packed)
            iqn_ops.pack_tc(self.flat, self.packed_tc)

    def named_views(self):
        out, o = OrderedDict(), 0
        for name, shape in iqn_ops.PARAM_SPECS:
            n = 1
            for d in shape:
                n *= d
            out[name] = self.flat[o:o + n].view(shape)
            o += n
        return out

    def state_dict(self):
        return OrderedDict((k, v.detach().clone()) for k, v in self.named_views().items())

    def load_state_dict(self, sd):
        views = self.named_views()
        missing = [k for k in views if k not in sd]
        if missing:
            raise KeyError(f"missing keys in state_dict: {missing}")
        for k, v in views.items():
            v.copy_(torch.as_tensor(sd[k]).to(self.device, torch.float32))
        self.repack()

    def parameters(self):
        return list(self.named_views().values())

    def to(self, device):
        if torch.device(device) != self.device:
            self.device = torch.device(device)
            self.flat = self.flat.to(self.device); self.packed = self.packed.to(self.device)
            self.packed_tc = self.packed_tc.to(self.device)
        return self

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        self.training = mode
        return self

    # ---- compute ---------------------------------------------------------------------------------------------------
    def draw_taus(self, batch_size, n_tau):
        """torch.rand(batch, n_tau) on the CPU generator then .to(device) (model.py:149), or a device generator if one was set."""
        if self.tau_generator is not None:
            return torch.rand(batch_size, n_tau, device=self.device, generator=self.tau_generator)
        return torch.rand(batch_size, n_tau).to(self.device)

    def forward(self, inputs, num_tau=8, cvar=1.0, taus=None):
        """-> (quantiles [B, num_tau, 9], taus [B, num_tau, 1]) like model.py:160-186; taus are returned already * cvar."""
        assert inputs.shape[1] == self.state_size, "input size not equal state size"
        x = inputs.to(self.device, torch.float32).contiguous()
        B = x.shape[0]
        if taus is None:
            taus = self.draw_taus(B, num_tau)
        with torch.cuda.device(self.device):
            q, _, _ = iqn_ops.forward(self.flat, self.packed, x, taus.contiguous(), cvar)
        cv = cvar.view(B, 1) if torch.is_tensor(cvar) else cvar
        return q, (taus * cv).unsqueeze(-1)

    __call__ = forward

    def get_qvals(self, inputs, cvar, taus=None):
        """Mean over K = 32 quantile samples (model.py:188-191)."""
        x = inputs.to(self.device, torch.float32).contiguous()
        if taus is None:
            taus = self.draw_taus(x.shape[0], self.K)
        with torch.cuda.device(self.device):
            _, qm, _ = iqn_ops.forward(self.flat, self.packed, x, taus.contiguous(), cvar, want_quantiles=False, want_qmean=True)
        return qm

    # ---- persistence (model.py:193-225) ------------------------------------------------------------------------------
    def get_constructor_parameters(self):
        return dict(state_size=self.state_size, action_size=self.action_size, seed=self.seed_id)

    def save(self, directory):
        torch.save(OrderedDict((k, v.cpu()) for k, v in self.state_dict().items()), os.path.join(directory, "network_params.pth"))
        with open(os.path.join(directory, "constructor_params.json"), mode="w") as f:
            json.dump(self.get_constructor_parameters(), f)

    @classmethod
    def load(cls, directory, device="cuda:0"):
        params = torch.load(os.path.join(directory, "network_params.pth"), map_location="cpu")
        with open(os.path.join(directory, "constructor_params.json"), mode="r") as f:
            ctor = json.load(f)
        ctor["device"] = device
        model = cls(**ctor)
        model.load_state_dict(params)
        return model
