"""Model layer: the mirror of gordo/machine/model/** (the drop-in boundary)."""
