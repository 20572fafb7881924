"""Action → force models that run immediately before ``World.step`` (ref vmas/simulator/dynamics/)."""
