from . import feedforward_autoencoder, lstm_autoencoder

__all__ = ["feedforward_autoencoder", "lstm_autoencoder"]
