"""Mirror of the reference package ``score_sde_pytorch`` (hot-path modules only)."""
