"""Mirror of the reference's ``modules`` package for the hot path: ``rendering`` and ``metrics``."""
