"""Drop-in for `code/centroidal_mpc_vertices_payload.py`: identical to the vertices module except for the
change-of-coordinates gains k1, k2 = 7, 1 (payload file :27-31; k2 cancels out of the NLP)."""
from .centroidal_mpc_vertices import centroidal_mpc as _Base


class centroidal_mpc(_Base):  # noqa: N801
    K1K2 = (7.0, 1.0)
