"""Mirror of captioning/models/crnn_trm_encoder.py:179-211 `CrnnEncoder` (HF copy hf_wrapper.py:1350-1374
`Cnn14RnnEncoder`): CNN encoder -> rename attn_emb / attn_emb_len -> RNN encoder."""
import torch.nn as nn


class CrnnEncoder(nn.Module):

    def __init__(self, cnn, rnn, freeze_cnn=False, freeze_cnn_bn=False, **kwargs):
        super().__init__()
        self.cnn = cnn
        self.rnn = rnn
        if freeze_cnn:
            for param in self.cnn.parameters():
                param.requires_grad = False
            self.freeze_cnn_bn = freeze_cnn_bn

    def forward(self, input_dict):
        output_dict = self.cnn(input_dict)
        output_dict["attn"] = output_dict["attn_emb"]
        output_dict["attn_len"] = output_dict["attn_emb_len"]
        del output_dict["attn_emb"], output_dict["attn_emb_len"]
        return self.rnn(output_dict)
