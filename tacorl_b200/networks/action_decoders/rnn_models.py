"""Mirror of /root/reference/src/tacorl/networks/action_decoders/rnn_models.py:5-16 (rnn_decoder only;
the lstm/gru/mlp variants are not selected by any shipped config)."""
from ..layers import ReluRNN


def rnn_decoder(in_features: int, hidden_size: int, num_layers: int, policy_rnn_dropout_p: float):
    return ReluRNN(in_features, hidden_size, num_layers=num_layers, bidirectional=False,
                   dropout=policy_rnn_dropout_p)
