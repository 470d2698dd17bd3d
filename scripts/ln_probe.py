import os, sys, torch
sys.path.insert(0, os.getcwd())
from bench import synthetic_state_dict, VOCAB
from sonar_b200 import B200TextEncoderModel, SequenceBatch, sonar_text_encoder_config
dev = torch.device("cuda:0")
model = B200TextEncoderModel(sonar_text_encoder_config("basic", num_encoder_layers=2), synthetic_state_dict(dev, layers=2), dev)
ids = torch.randint(4, VOCAB, (4096, 128), device=dev)
for _ in range(2):
    model(SequenceBatch(ids, None))
torch.cuda.synchronize()
