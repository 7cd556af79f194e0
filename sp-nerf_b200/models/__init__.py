"""Mirror of the reference's ``models`` package surface (models/__init__.py:1-16)."""
from .spnerf import SPNeRF, Sine, PositionalEncoding, inference  # noqa: F401


def load_model(args):
    """Factory with the reference's contract (models/__init__.py:4-16): reads
    num_sem_classes, s_embedding_factor, fc_layers, fc_units, mapping, t_embbeding_tau, beta, sem."""
    if args.model != "sp-nerf":
        raise ValueError(f'model {args.model} is not valid')
    return SPNeRF(num_sem_classes=args.num_sem_classes, s_embedding_factor=args.s_embedding_factor,
                  layers=args.fc_layers, feat=args.fc_units, mapping=args.mapping,
                  t_embedding_dims=args.t_embbeding_tau, beta=args.beta, sem=args.sem)
