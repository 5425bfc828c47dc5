"""rsoccer_b200 -- B200-native batched robot-soccer simulator (rSoccer-compatible).

Replaces the physics back-end of robocin/rSoccer (rsoccer_gym/Simulators -> robosim) with
hand-written sm_100a CUDA kernels stepping N independent matches in lockstep.
"""
__version__ = "0.1.0"
