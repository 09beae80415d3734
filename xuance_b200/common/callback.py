"""Training callback hooks (reference: xuance/common/callback.py:4-79).  Every hook is a no-op returning an
empty dict (or None) so learners/agents can call them unconditionally, as the reference does."""


class BaseCallback:
    def on_update_start(self, *args, **kwargs):
        return {}

    def on_update_end(self, *args, **kwargs):
        return {}

    def on_update_agent_wise(self, *args, **kwargs):
        return {}

    def on_train_step(self, *args, **kwargs):
        return None

    def on_train_epochs_end(self, *args, **kwargs):
        return None

    def on_train_episode_info(self, *args, **kwargs):
        return None

    def on_train_step_end(self, *args, **kwargs):
        return None

    def on_test_step(self, *args, **kwargs):
        return None

    def on_test_end(self, *args, **kwargs):
        return None
