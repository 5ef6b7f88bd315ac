"""infinitensor_b200 -- B200-native operator-kernel backend behind InfiniTensor's
GraphObj / RuntimeObj / KernelRegistry contract.  See DESIGN.md."""
__all__ = ["backend"]
