"""Host logic of GanTrainer.run_reference_schedule against a literal transcription of the reference's loop control flow
(main.py:137-172): which batches trigger a discriminator Nadam step, which trigger generator passes, how many."""
import torch

from locate_b200.train import GanTrainer


class _Recorder(GanTrainer):
    def __init__(self):
        self.log = []
        self.d_opt, self.g_opt = "D", "G"

    def d_step(self, real, aug, z, update=True):
        self.log.append(("d_backward", int(real[0])))
        assert not update
        return "d"

    def g_step(self, z, update=True):
        self.log.append(("g_backward",))
        assert not update
        return "g"

    def _reduce_and_step(self, opt):
        self.log.append(("step", opt))


def _reference_control_flow(n_batches, miniter, minibatches, diters):
    log = []
    for i in range(1, n_batches + 1):
        log.append(("d_backward", i))                 # dis.zero_grad(); ...; (d_error + penalty).backward()
        if i % miniter == 0:
            log.append(("step", "D"))                 # DIS_OPTIM.step()
            if (i // miniter) % diters == 0:
                for _ in range(minibatches):
                    log.append(("g_backward",))       # gen.zero_grad(); g_error.backward()
                log.append(("step", "G"))             # GEN_OPTIM.step()
    return log


def test_schedule_matches_reference_control_flow():
    for miniter, minibatches, diters in [(1, 1, 1), (8, 8, 1), (2, 3, 2), (3, 1, 4)]:
        tr = _Recorder()
        batches = [(torch.tensor([i]), torch.tensor([i])) for i in range(1, 26)]
        out = list(tr.run_reference_schedule(batches, miniter, minibatches, diters, noise_fn=lambda n: torch.zeros(n, 4)))
        assert tr.log == _reference_control_flow(25, miniter, minibatches, diters), (miniter, minibatches, diters)
        assert [o[0] for o in out] == list(range(1, 26))
        g_steps = [o[2] for o in out if o[2] is not None]
        assert len(g_steps) == sum(1 for e in tr.log if e == ("step", "G"))
