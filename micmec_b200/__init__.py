"""micmec_b200: B200-native force evaluation + MD integration for the micromechanical model (MicMec).

Only the hot path of molmod/micmec lives here: ``pes.mmff`` (ForcePartMechanical behind the ForcePart plugin API)
and ``sampling`` (VerletIntegrator with device-resident NHC thermostat / MTK barostat).  The arithmetic runs in
hand-written fp64 CUDA kernels (``csrc/``, sm_100a) behind the C ABI of ``include/micmec_b200.h``.
"""
__version__ = "0.1.0"
