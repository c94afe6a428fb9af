"""helper_interfaces package of the B200 density-matrix engine repository."""
