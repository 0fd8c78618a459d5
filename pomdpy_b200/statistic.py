"""Running statistics of an experiment (mirrors reference pomdpy/pomdp/statistic.py:9-47)."""
import math


class Statistic(object):
    def __init__(self, name, val=0.0, count=0.0):
        self.name = name
        self.mean = val
        self.count = count
        self.variance = 0.0
        self.max = -math.inf
        self.min = math.inf
        self.running_total = 0.0

    def add(self, val):
        """Incremental mean / population variance, exactly the reference's update (statistic.py:21-33)."""
        self.running_total += val
        prev_mean, prev_count = self.mean, self.count
        self.count += 1
        self.mean += (val - self.mean) / self.count
        second_moment = prev_count * (self.variance + prev_mean * prev_mean) + val * val
        self.variance = second_moment / self.count - self.mean * self.mean
        self.max = max(self.max, val)
        self.min = min(self.min, val)

    def std_dev(self):
        return math.sqrt(max(self.variance, 0.0))

    def std_err(self):
        return math.sqrt(max(self.variance, 0.0) / self.count)

    def clear(self):
        self.mean, self.count, self.variance = 0.0, 0, 0.0
        self.max, self.min = -math.inf, math.inf

    def show(self):
        for label, v in (("Name", self.name), ("Running Total", self.running_total), ("Mean", self.mean),
                         ("Count", self.count), ("Variance", self.variance), ("Max", self.max), ("Min", self.min)):
            print(label + " = ", v)
