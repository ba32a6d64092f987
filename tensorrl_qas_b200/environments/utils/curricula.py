"""Threshold curricula consulted by CircuitEnv (reference: environments/utils/curricula.py).  Pure bookkeeping above
the kernel boundary; same class names, constructor signature `(config, target_energy=...)`, attributes
(`lowest_energy`, `current_threshold`) and methods (`get_current_threshold`, `update_threshold(energy_done=...)`)."""


class VanillaCurriculum:
    """Piecewise-constant schedule: thresholds[i] applies until switch_episodes[i] episodes have finished
    (curricula.py:79-97).  The only curriculum the shipped cfgs use."""

    def __init__(self, config, **kw):
        self.thresholds = config["thresholds"]
        self.episodes = config["switch_episodes"]
        self.episodes_completed = 0
        self.min_en = kw.get("target_energy")
        self.current_threshold = config["accept_err"]
        self.lowest_energy = self.min_en + self.current_threshold

    def get_current_threshold(self):
        pending = [i for i, limit in enumerate(self.episodes) if limit > self.episodes_completed]
        return self.thresholds[min(pending)]  # ValueError once every switch point has passed, as in the reference

    def update_threshold(self, **kw):
        self.episodes_completed += 1


class MovingThreshold:
    """Threshold that follows the lowest energy seen, with an amortisation radius (curricula.py:1-53)."""

    def __init__(self, config, **kw):
        self.amortisation = config["shift_threshold_ball"]
        self.greedy_shift_time = config["shift_threshold_time"]
        self.min_en = kw.get("target_energy")
        self.success_thresh = config["success_thresh"]
        self.succ_radius_shift = config["succ_radius_shift"]
        self.succes_switch = config["succes_switch"]
        self.current_threshold = config["accept_err"]
        self.lowest_energy = self.min_en + self.current_threshold
        self.success_counter = 0
        self.radius_shift_counter = 0
        self.call_counter = 0

    def _gap(self):
        return abs(self.min_en - self.lowest_energy)

    def reduce_amortisation(self):
        if self.success_thresh:
            self.success_counter += 1
            ripe = self.success_counter >= self.success_thresh
            if ripe and self.radius_shift_counter < self.succ_radius_shift and self.succes_switch > self._gap():
                self.current_threshold -= self.amortisation / self.succ_radius_shift
                self.success_counter = 0
                self.radius_shift_counter += 1
        return self.current_threshold

    def greedy_shift(self):
        self.call_counter += 1
        if self.call_counter > 10 and self.call_counter % self.greedy_shift_time == 0:
            self.current_threshold = self._gap() + (self.amortisation if self.amortisation else 0)
            if self.amortisation and self.success_thresh:
                self.radius_shift_counter = 0
                self.success_counter = 0
        return self.current_threshold

    def get_current_threshold(self):
        return self.current_threshold

    def update_threshold(self, **kw):
        if kw.get("energy_done"):
            self.reduce_amortisation()
        self.greedy_shift()


class SuccesCountThreshold:
    """Threshold tightened to the best gap after `success_thresh` successes (curricula.py:55-77)."""

    def __init__(self, config, **kw):
        self.min_en = kw.get("target_energy")
        self.success_thresh = config["success_thresh"]
        self.current_threshold = config["accept_err"]
        self.lowest_energy = self.min_en + self.current_threshold
        self.success_counter = 0

    def greedy_shift(self):
        if self.success_thresh:
            self.success_counter += 1
            if self.success_counter >= self.success_thresh:
                self.success_counter = 0
                self.current_threshold = abs(self.min_en - self.lowest_energy)
        return self.current_threshold

    def get_current_threshold(self):
        return self.current_threshold

    def update_threshold(self, **kw):
        if kw.get("energy_done"):
            self.greedy_shift()
