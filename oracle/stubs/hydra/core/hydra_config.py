class HydraConfig:
    @staticmethod
    def get():
        raise RuntimeError("hydra stub: no HydraConfig")
