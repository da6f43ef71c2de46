"""Mirror of captioning/models/crnn_trm_encoder.py:179-211 `CrnnEncoder` (HF copy hf_wrapper.py:1350-1374
`Cnn14RnnEncoder`): CNN encoder -> rename attn_emb / attn_emb_len -> RNN encoder."""
import torch.nn as nn


class CrnnEncoder(nn.Module):

    def __init__(self, cnn, rnn, freeze_cnn=False, freeze_cnn_bn=False, **kwargs):
        super().__init__()
        self.cnn = cnn
        self.rnn = rnn
        self.freeze_cnn_bn = False
        if freeze_cnn:
            for param in self.cnn.parameters():
                param.requires_grad = False
            self.freeze_cnn_bn = freeze_cnn_bn

    def train(self, mode=True):
        """crnn_trm_encoder.py:195-203: with `freeze_cnn_bn` every *BatchNorm* module of the CNN stays in eval mode
        while the rest of the encoder trains (dropout inside the frozen CNN stays active)."""
        super().train(mode)
        if self.freeze_cnn_bn:
            for module in self.cnn.modules():
                if "BatchNorm" in module.__class__.__name__:
                    module.eval()
        return self

    def forward(self, input_dict):
        output_dict = self.cnn(input_dict)
        output_dict["attn"] = output_dict["attn_emb"]
        output_dict["attn_len"] = output_dict["attn_emb_len"]
        del output_dict["attn_emb"], output_dict["attn_emb_len"]
        return self.rnn(output_dict)


# eg_configs/clotho_v2/waveform/cnn14rnn_trm.yaml:9 names this class; the reference file does not define it (the HF
# copy hf_wrapper.py:1350-1374 does, with the same body), so the Clotho YAML only resolves here.
Cnn14RnnEncoder = CrnnEncoder
