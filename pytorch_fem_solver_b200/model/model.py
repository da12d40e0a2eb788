"""Adam training driver (reference torch_fem/model/model.py) -- orchestration only."""

from __future__ import annotations

from typing import Callable, Optional

import torch


class Model:
    """zero_grad -> training_step -> backward -> step, with history and best-state tracking."""

    def __init__(
        self,
        neural_network: torch.nn.Module,
        training_step: Callable,
        epochs: int = 5000,
        optimizer: type[torch.optim.Optimizer] = torch.optim.Adam,
        optimizer_kwargs: Optional[dict] = None,
        learning_rate_scheduler=None,
        scheduler_kwargs: Optional[dict] = None,
        use_early_stopping: bool = False,
        early_stopping_patience: int = 10,
        min_delta: float = 1e-12,
        progress: bool = True,
    ):
        self._neural_network = neural_network
        self._training_step = training_step
        self._epochs = epochs
        self._optimizer = optimizer(self._neural_network.parameters(), **(optimizer_kwargs or {"lr": 0.001}))
        self._learning_rate_scheduler = (
            learning_rate_scheduler(self._optimizer, **(scheduler_kwargs or {})) if learning_rate_scheduler else None
        )
        self._use_early_stopping = use_early_stopping
        self._early_stopping_patience = early_stopping_patience
        self._min_delta = min_delta
        self._loss_history: list[float] = []
        self._validation_loss_history: list[float] = []
        self._accuracy_history: list[float] = []
        self._progress = progress
        self._best_loss = float("inf")
        self.optimal_parameters = {k: v.clone() for k, v in self._neural_network.state_dict().items()}
        self.early_stopping_counter = 0

    def _iterator(self):
        if not self._progress:
            return range(self._epochs), None
        try:
            import tqdm

            bar = tqdm.tqdm(range(self._epochs), desc="Training Progress")
            return bar, bar
        except ImportError:  # pragma: no cover
            return range(self._epochs), None

    def train(self):
        iterator, bar = self._iterator()
        for _ in iterator:
            self._optimizer.zero_grad()
            loss, validation_loss, accuracy = self._training_step(self._neural_network)
            loss.backward()
            self._optimizer.step()
            if self._learning_rate_scheduler is not None:
                self._learning_rate_scheduler.step(loss.detach())
            loss_value, validation_value, accuracy_value = loss.item(), validation_loss.item(), accuracy.item()
            threshold = self._best_loss - (self._min_delta if self._use_early_stopping else 0.0)
            if loss_value < threshold:
                self._best_loss = loss_value
                self.early_stopping_counter = 0
                self.optimal_parameters = {k: v.clone() for k, v in self._neural_network.state_dict().items()}
            elif self._use_early_stopping:
                self.early_stopping_counter += 1
                if self.early_stopping_counter >= self._early_stopping_patience:
                    break
            if bar is not None:
                bar.set_postfix(
                    {"Loss": f"{loss_value:.8f}", "Validation loss": f"{validation_value:.8f}", "Accuracy": f"{accuracy_value:.8f}"}
                )
            self._loss_history.append(loss_value)
            self._validation_loss_history.append(validation_value)
            self._accuracy_history.append(accuracy_value)

    def get_training_history(self):
        return self._loss_history, self._validation_loss_history, self._accuracy_history

    def load_optimal_parameters(self):
        self._neural_network.load_state_dict(self.optimal_parameters)

    def plot_training_history(self, plot_names: Optional[dict] = None):
        """Semilog plot of the three histories; needs matplotlib (not part of this image)."""
        import matplotlib.pyplot as plt

        names = plot_names or {"loss": "Training loss", "validation": "Validation loss", "accuracy": "Accuracy", "title": "Training history"}
        _, axis = plt.subplots()
        axis.semilogy(self._loss_history, linestyle="-", label=names["loss"])
        axis.semilogy(self._validation_loss_history, linestyle="--", label=names["validation"])
        axis.semilogy(self._accuracy_history, linestyle=":", label=names["accuracy"])
        axis.set_xlabel("# Epochs")
        axis.set_ylabel("Loss")
        axis.set_title(names["title"])
        axis.legend()
