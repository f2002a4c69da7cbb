"""Mass / ForceField path for N instances with device-resident tiled-SoA buffers (torch = allocator + streams).

``BatchedDynamics.eval`` is the BASELINE metric unit: forward kinematics + world joint Jacobian + CRBA mass
matrix + RNEA torques in ONE fused kernel launch (``rbd_eval_soa``).  Layout of every tensor: shape
``[tiles, K, tile]`` float64 (``tile == n`` gives the plain SoA ``[1, K, n]``), see ``include/rbd_b200.h``.
"""
from __future__ import annotations

from typing import Dict, Optional

from .api import Model, Workspace
from .model import ModelDescriptor


class BatchedDynamics:
    def __init__(self, desc: ModelDescriptor, n: int, device: int = 0, tile: Optional[int] = None):
        import torch
        self.torch = torch
        self.desc = desc
        self.n = int(n)
        self.tile = int(tile) if tile else self.n
        self.ntiles = (self.n + self.tile - 1) // self.tile
        self.device = torch.device("cuda", device)
        self.model = Model(desc)
        self.ws = Workspace(self.model, self.n, device)
        I = self.model.info
        self.fields = {"q": I.nq, "v": I.nv, "a": I.nv, "oMi": I.omi_fields, "J": I.jac_fields, "M": I.mass_fields,
                       "tau": I.nv}

    def empty(self, name: str):
        return self.torch.empty((self.ntiles, self.fields[name], self.tile), dtype=self.torch.float64, device=self.device)

    def alloc_outputs(self) -> Dict[str, "object"]:
        return {k: self.empty(k) for k in ("oMi", "J", "M", "tau")}

    def to_soa(self, x_instance_major):
        """[n, K] (any device) -> tiled-SoA device tensor, through the library's layout adapter."""
        torch = self.torch
        x = x_instance_major.to(self.device, torch.float64).contiguous()
        n, K = x.shape
        out = torch.zeros((self.ntiles, K, self.tile), dtype=torch.float64, device=self.device)
        self._on_current_stream()
        self.ws.aos_to_soa(n, K, self.tile, x, out)
        return out

    def to_instance_major(self, x_soa, n: Optional[int] = None):
        torch = self.torch
        n = self.n if n is None else n
        K = x_soa.shape[1]
        out = torch.empty((n, K), dtype=torch.float64, device=self.device)
        self._on_current_stream()
        self.ws.soa_to_aos(n, K, self.tile, x_soa, out)
        return out

    def _on_current_stream(self):
        self.ws.set_stream(self.torch.cuda.current_stream(self.device).cuda_stream)

    def eval(self, q, v, a, out: Dict[str, "object"]):
        """One fused launch; any of out['oMi'|'J'|'M'|'tau'] may be missing/None."""
        self._on_current_stream()
        self.ws.eval_soa(self.n, self.tile, q, v, a, out.get("oMi"), out.get("J"), out.get("M"), out.get("tau"))

    def crba(self, q, M):
        self._on_current_stream()
        self.ws.crba_soa(self.n, self.tile, q, M)

    def rnea(self, q, v, a, tau):
        self._on_current_stream()
        self.ws.rnea_soa(self.n, self.tile, q, v, a, tau)

    # -- SOFA Mass / ForceField interface in joint space (reference: UniformMass<Rigid3> per body hung under the
    #    mapping, URDFModelLoader.cpp:369-391; reached there matrix-free through applyJ / applyJT) -------------
    def addMDx(self, q, dx, factor, res):
        """Mass::addMDx: res += factor * M(q) dx, one RNEA-shaped sweep (no M is formed)."""
        self._on_current_stream()
        self.ws.addMDx_soa(self.n, self.tile, q, dx, factor, res)

    def addForce(self, q, v, f):
        """ForceField::addForce: f -= C(q, v) v + g(q); v None = gravity only (what the reference computes)."""
        self._on_current_stream()
        self.ws.addForce_soa(self.n, self.tile, q, v, f)
