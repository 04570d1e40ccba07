"""B200-native 2-D Barnes-Hut engine: drop-in for the simulation path of
DavidSevic/gpu-nbody-simulation (implementation/project.cu).  See DESIGN.md."""
