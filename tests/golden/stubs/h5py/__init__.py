"""Empty stand-in: only ``import h5py`` itself is needed on the force/MD path."""
