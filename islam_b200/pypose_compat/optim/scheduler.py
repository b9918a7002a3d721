"""pypose.optim.scheduler.StopOnPlateau as used at /root/reference/pvgo.py:172-180."""


class StopOnPlateau:
    def __init__(self, optimizer, steps, patience=5, decreasing=1e-3, verbose=False):
        self.optimizer, self.max_steps, self.patience = optimizer, steps, patience
        self.decreasing, self.verbose = decreasing, verbose
        self.steps, self.patience_count, self._continual = 0, 0, True

    def continual(self):
        return self._continual

    def step(self, loss):
        assert self.optimizer.loss is not None, 'scheduler.step() should be called after optimizer.step()'
        if self.verbose:
            print('StopOnPlateau on step {} Loss {:.6e} --> Loss {:.6e} (reduction/loss: {:.4e}).'.format(
                self.steps, float(self.optimizer.last), float(self.optimizer.loss),
                float((self.optimizer.last - self.optimizer.loss) / (self.optimizer.last + 1e-31))))
        self.steps += 1
        if self.steps >= self.max_steps:
            self._continual = False
        if float(self.optimizer.last - loss) < self.decreasing:
            self.patience_count += 1
        else:
            self.patience_count = 0
        if self.patience_count >= self.patience:
            self._continual = False
        if hasattr(self.optimizer, 'reject') and self.optimizer.reject_count >= self.optimizer.reject:
            self._continual = False
