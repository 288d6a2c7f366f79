// spvopt.cpp — TEST INFRASTRUCTURE: puts a SPIR-V module into the form the reference's pipeline holds it in.
//
// vk::GraphicsPipeline::compileShaders runs every module through spirv-opt before sw::SpirvShader sees it
// (/root/reference/src/Vulkan/VkPipeline.cpp:36-107: CreateRemoveDontInlinePass + RegisterPerformancePasses, validator off in release
// builds), and `SpirvShader::insns` — what the shim hands to swcu_shader_translate (icd/swcu_shim.cpp) — is that optimised binary.
// This tool applies the same pass list with the SPIRV-Tools of the reference build (oracle/build_ref.sh leaves the static libraries in
// $SS_BUILD_DIR), so that the translator can be tested on both forms without a GPU (tests/test_boundary.py, fixtures under
// tests/golden/spv_postopt/).  Built by oracle/Makefile (target _ref/spvopt) only where /root/reference and the build tree exist.
//
// usage: spvopt in.spv out.spv [--dis]
#include "spirv-tools/libspirv.hpp"
#include "spirv-tools/optimizer.hpp"

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

int main(int argc, char **argv)
{
	if(argc < 3) { fprintf(stderr, "usage: %s in.spv out.spv [--dis]\n", argv[0]); return 2; }
	FILE *f = fopen(argv[1], "rb");
	if(!f) { perror(argv[1]); return 1; }
	std::vector<uint32_t> code;
	uint32_t w;
	while(fread(&w, 4, 1, f) == 1) code.push_back(w);
	fclose(f);

	const spv_target_env env = SPV_ENV_VULKAN_1_3; // vk::SPIRV_VERSION (VkConfig.hpp)
	spvtools::Optimizer opt{ env };
	opt.SetMessageConsumer([](spv_message_level_t, const char *, const spv_position_t &p, const char *m) { fprintf(stderr, "spirv-opt: %d:%d %s\n", (int)p.line, (int)p.column, m); });
	opt.RegisterPass(spvtools::CreateRemoveDontInlinePass());
	opt.RegisterPerformancePasses();
	spvtools::OptimizerOptions options = {};
	options.set_run_validator(false); // NDEBUG build of the reference
	std::vector<uint32_t> out;
	if(!opt.Run(code.data(), code.size(), &out, options) || out.empty()) { fprintf(stderr, "spirv-opt failed\n"); return 1; }
	f = fopen(argv[2], "wb");
	if(!f) { perror(argv[2]); return 1; }
	fwrite(out.data(), 4, out.size(), f);
	fclose(f);
	if(argc > 3 && std::string(argv[3]) == "--dis")
	{
		spvtools::SpirvTools core(env);
		std::string text;
		core.Disassemble(out, &text, SPV_BINARY_TO_TEXT_OPTION_FRIENDLY_NAMES | SPV_BINARY_TO_TEXT_OPTION_INDENT);
		puts(text.c_str());
	}
	return 0;
}
