python tools/emd_time.py base
PCL_EMD_PCAP=64 PCL_EMD_ITEMS=48 python tools/emd_time.py items48
PCL_EMD_PCAP=64 PCL_EMD_ITEMS=64 python tools/emd_time.py items64
PCL_EMD_PCAP=96 PCL_EMD_ITEMS=96 python tools/emd_time.py items96
PCL_EMD_PCAP=128 PCL_EMD_ITEMS=128 python tools/emd_time.py items128
PCL_EMD_PCAP=64 PCL_EMD_ITEMS=64 PCL_EMD_WPB=64 python tools/emd_time.py items64_wpb64
PCL_EMD_PCAP=64 PCL_EMD_ITEMS=64 PCL_EMD_WPB=128 python tools/emd_time.py items64_wpb128
