"""B200-native batched articulated-body dynamics back end for the Sofa.RigidBodyDynamics plugin.

Scope = the one hot path of the reference: the pinocchio work behind
``KinematicChainMapping::apply / applyJ / applyJT`` and the CRBA / RNEA quantities of the Mass/ForceField path,
for many robot instances at once, as hand-written sm_100a CUDA kernels behind the C ABI in ``include/rbd_b200.h``.

Importing this package does not need a GPU; loading ``librbd_b200.so`` does need the library to have been built
(``python -m sofa.rigidbodydynamics_b200.build``), and every compute call needs a CUDA device — there is no CPU
or PyTorch fallback.
"""
from . import model
from .model import ModelDescriptor, from_urdf, from_urdf_string, panda7, solo12, talos_reduced, random_tree

__all__ = ["model", "ModelDescriptor", "from_urdf", "from_urdf_string", "panda7", "solo12", "talos_reduced",
           "random_tree", "Model", "Workspace", "KinematicChainMapping", "JointSpaceMass", "BatchedDynamics", "shard_range"]


def __getattr__(name):  # lazy: keep `import sofa.rigidbodydynamics_b200` free of ctypes loading
    if name in ("Model", "Workspace", "RbdError", "device_count"):
        from . import api
        return getattr(api, name)
    if name == "KinematicChainMapping":
        from .mapping import KinematicChainMapping
        return KinematicChainMapping
    if name == "JointSpaceMass":
        from .mass import JointSpaceMass
        return JointSpaceMass
    if name in ("BatchedDynamics",):
        from .dynamics import BatchedDynamics
        return BatchedDynamics
    if name in ("shard_range",):
        from .sharding import shard_range
        return shard_range
    raise AttributeError(name)
