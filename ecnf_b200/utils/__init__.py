from .optim import Adam, AdamState, warmup_cosine_decay_schedule  # noqa: F401
