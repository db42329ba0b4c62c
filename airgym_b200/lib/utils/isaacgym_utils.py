"""RLGPUAlgoObserver — the observer object scripts/runner.py hands to Runner (lib/utils/isaacgym_utils.py:51-110).

In the reference it keeps a host-side list of every step's `item_reward_info` dict and averages it after each epoch into the
`Episode/<reward term>` TensorBoard scalars.  Here the agent sums the reward-term planes on the device inside the rollout graph
(A2CAgent._observe_reward_terms) and writes the same tags (A2CAgent.write_stats); this class keeps the reference's interface so
that `Runner(RLGPUAlgoObserver())` works unchanged, and forwards direct scalar infos."""
import torch


class AlgoObserver:
    def before_init(self, base_name, config, experiment_name):
        pass

    def after_init(self, algo):
        pass

    def process_infos(self, infos, done_indices):
        pass

    def after_steps(self):
        pass

    def after_clear_stats(self):
        pass

    def after_print_stats(self, frame, epoch_num, total_time):
        pass


class RLGPUAlgoObserver(AlgoObserver):
    def __init__(self):
        self.algo, self.writer, self.direct_info = None, None, {}

    def after_init(self, algo):
        self.algo, self.writer = algo, algo.writer

    def process_infos(self, infos, done_indices):
        assert isinstance(infos, dict), "RLGPUAlgoObserver expects dict info"
        self.direct_info = {k: v for k, v in infos.items()
                            if isinstance(v, (float, int)) or (isinstance(v, torch.Tensor) and v.dim() == 0)}

    def after_print_stats(self, frame, epoch_num, total_time):
        if self.writer is None:
            return
        for k, v in self.direct_info.items():
            self.writer.add_scalar(f"{k}/frame", v, frame)
            self.writer.add_scalar(f"{k}/iter", v, epoch_num)
            self.writer.add_scalar(f"{k}/time", v, total_time)
