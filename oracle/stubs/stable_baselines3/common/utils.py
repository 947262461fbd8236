def set_random_seed(seed):
    pass
