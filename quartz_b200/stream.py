"""Block-rate streaming adapter: the offline analogue of the reference's cpal pull loop (src/audio.rs:77-118).

The reference wraps the current graph in `BlockRateAdapter(SlotBackend)` and pulls one stereo frame per device frame,
sanitising (non-normal -> 0) and clamping each sample before it is interleaved into the device buffer.  Here the graph
lives on the GPU: `StreamAdapter` keeps a one-voice bank, renders look-ahead blocks with `qg_bank_render_stereo` and
serves frames from them; `set(net)` swaps the graph (`slot.set`, src/process.rs:1897), `set_var` forwards `var()`
updates (src/process.rs:1382-1385) which take effect at the next block boundary."""
import numpy as np

from .net import Bank, Net


class StreamAdapter:
    def __init__(self, net=None, block=4096, ctx=None):
        self.block, self.ctx = int(block), ctx
        self._buf = np.zeros((0, 2), np.float32)
        self._pos = 0
        self.bank = None
        self.set(net if net is not None else Net.str_to_net("dc(0)") | Net.str_to_net("dc(0)"))

    def set(self, net):
        """`out()` semantics (src/process.rs:1893-1905): 0-in/1-out -> net | dc(0); 0-in/2-out -> net; else silence."""
        if net.inputs() == 0 and net.outputs() in (1, 2):
            self.net = net
        else:
            self.net = Net.str_to_net("dc(0)") | Net.str_to_net("dc(0)")
        self.bank = Bank(self.net, 1, ctx=self.ctx)
        self._buf, self._pos = np.zeros((0, 2), np.float32), 0

    def set_var(self, raw_index, value):
        self.bank.set_raw(raw_index, value)

    def read(self, n_frames):
        """next n interleaved stereo frames [n, 2] (what `write_data` copies into the device buffer)"""
        out = np.empty((n_frames, 2), np.float32)
        done = 0
        while done < n_frames:
            if self._pos >= len(self._buf):
                self._buf, self._pos = self.bank.render_stereo(self.block), 0
            k = min(n_frames - done, len(self._buf) - self._pos)
            out[done:done + k] = self._buf[self._pos:self._pos + k]
            self._pos += k
            done += k
        return out
