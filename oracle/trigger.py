"""Oracle restatement of the activation debouncer.  TEST INFRASTRUCTURE ONLY.  Pinned by import.

Follows reference ``runner/precise_runner/runner.py:115-142`` (class ``TriggerDetector``).
``tests/golden/trigger_golden.npz`` holds traces produced by the reference class itself.
"""


class OracleTrigger:
    def __init__(self, chunk_size, sensitivity=0.5, trigger_level=3):
        self.chunk_size = chunk_size            # BYTES per prediction (runner.py:44-45)
        self.sensitivity = sensitivity
        self.trigger_level = trigger_level
        self.activation = 0

    def update(self, prob: float) -> bool:
        hot = prob > 1.0 - self.sensitivity
        if hot or self.activation < 0:
            self.activation += 1
            fired = self.activation > self.trigger_level
            # python precedence: fired or (hot and activation < 0)
            if fired or (hot and self.activation < 0):
                self.activation = -(8 * 2048) // self.chunk_size
            if fired:
                return True
        elif self.activation > 0:
            self.activation -= 1
        return False
