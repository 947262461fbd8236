class WandbLogger:
    pass
