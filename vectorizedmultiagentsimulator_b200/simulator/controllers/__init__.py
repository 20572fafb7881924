"""Action pre-processors scenarios call from ``process_action`` (ref vmas/simulator/controllers/)."""
