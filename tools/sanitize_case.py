"""Small end-to-end run for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dai_b200
from dai_b200 import synthetic
from dai_b200.torchmodel import ActiveInferenceModel
m = ActiveInferenceModel(10, 4, 1.0, 1.0, 1.0).load_numpy_weights(synthetic.make_weights(0))
o = torch.from_numpy(synthetic.make_frames(1, 1)).repeat(4, 1, 1, 1)
G, terms, po1 = m.calculate_G_4_repeated(o, steps=2, samples=3)
g2 = m.calculate_G_mean(torch.zeros(4, 10), torch.eye(4))[0]
g3 = m.mcts_step_simulate(torch.zeros(10), 3)[0]
c = m.select_actions(torch.from_numpy(synthetic.make_frames(2, 2)), steps=1, samples=2)[0]
g4 = m.mcts_step_simulate_batch(torch.zeros(3, 10), 3)[0]
from oracle import frames_oracle as F
from dai_b200.game_environment import FrameProducer
fp = FrameProducer(F.make_sprites((1, 3, 2, 2, 4, 4), 0), (1, 3, 2, 2, 4, 4), engine=m._engine)
fr = fp.current_frame_all(torch.tensor([[0, 1, 1, 1, 2, 3, 0.0], [0, 2, 0, 0, 1, 1, 0.0]]), torch.tensor([0.5, -0.25]))
torch.cuda.synchronize()
print("ok", G.tolist(), g2.tolist(), g3, c.tolist(), g4.tolist(), float(fr.sum()))
