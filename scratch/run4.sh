python profiles/trace_layer.py 288 16 > gpurun_out/trace_layer_v5.log 2>&1; cat gpurun_out/trace_layer_v5.log | tail -32
