// spirv_subset.cpp — the narrow SPIR-V front end of the B200 draw path (host only, no GPU needed).
//
// Stands in for sw::SpirvShader's analysis of the two graphics stages (reference:
// src/Pipeline/SpirvShader.cpp:947-961 interface slots, src/Pipeline/SpirvShader.hpp:761-782 decorations,
// src/Pipeline/VertexProgram.cpp:75-94, src/Pipeline/PixelProgram.cpp:138-241) for the benchmark subset ONLY:
//   vertex:   gl_Position and user varyings are built from vertex inputs, float constants and the push-constant block with copies /
//             swizzles and straight-line float arithmetic (OpFAdd / OpFSub / OpFMul / OpFNegate / OpVectorTimesScalar /
//             OpMatrixTimesVector / OpVectorTimesMatrix / OpMatrixTimesScalar / OpDot): an MVP transform.  The arithmetic is lowered to
//             a list of scalar steps with the reference's rounding (SpirvShaderArithmetic.cpp:39-75,449-457,611-621: products and
//             sums are single operations, matrix and dot products accumulate with MulAdd);
//   fragment: colour output 0 is built from interpolated inputs, float constants and at most one
//             OpImageSampleImplicitLod of a combined image sampler whose coordinate is again such a value.
// The result is an operand-routing table (swcu_shader_info); the kernels in kernels.cuh are specialised on it, so the
// "device function" a module lowers to is a fixed gather of registers.  Everything else is REJECTED
// (SWCU_E_UNSUPPORTED) — there is no interpreter and no CPU fallback.
#include "swcu_internal.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

namespace {

enum Op : uint32_t
{
	OpNop = 0, OpSourceContinued = 2, OpSource = 3, OpSourceExtension = 4, OpName = 5, OpMemberName = 6, OpString = 7, OpLine = 8,
	OpExtension = 10, OpExtInstImport = 11, OpMemoryModel = 14, OpEntryPoint = 15, OpExecutionMode = 16, OpCapability = 17,
	OpTypeVoid = 19, OpTypeBool = 20, OpTypeInt = 21, OpTypeFloat = 22, OpTypeVector = 23, OpTypeMatrix = 24, OpTypeImage = 25, OpTypeSampler = 26,
	OpTypeSampledImage = 27, OpTypeArray = 28, OpTypeStruct = 30, OpTypePointer = 32, OpTypeFunction = 33,
	OpConstant = 43, OpConstantComposite = 44, OpFunction = 54, OpFunctionEnd = 56, OpVariable = 59, OpLoad = 61, OpStore = 62,
	OpAccessChain = 65, OpInBoundsAccessChain = 66, OpDecorate = 71, OpMemberDecorate = 72, OpVectorShuffle = 79,
	OpCompositeConstruct = 80, OpCompositeExtract = 81, OpCompositeInsert = 82, OpCopyObject = 83,
	OpImageSampleImplicitLod = 87, OpFNegate = 127, OpFAdd = 129, OpFSub = 131, OpFMul = 133, OpVectorTimesScalar = 142,
	OpMatrixTimesScalar = 143, OpVectorTimesMatrix = 144, OpMatrixTimesVector = 145, OpDot = 148, OpLabel = 248, OpReturn = 253, OpNoLine = 317, OpModuleProcessed = 330,
};
enum { DecBlock = 2, DecRowMajor = 4, DecColMajor = 5, DecMatrixStride = 7, DecBuiltIn = 11, DecNoPerspective = 13, DecFlat = 14, DecLocation = 30, DecComponent = 31, DecBinding = 33,
	   DecDescriptorSet = 34, DecOffset = 35, DecRelaxedPrecision = 0 };
enum { BuiltInPosition = 0, BuiltInPointSize = 1, BuiltInClipDistance = 3, BuiltInCullDistance = 4 };
enum { SCUniformConstant = 0, SCInput = 1, SCUniform = 2, SCOutput = 3, SCPushConstant = 9 };

struct Type
{
	enum Kind { None, Void, Float, Int, Vector, Matrix, Struct, Pointer, Image, SampledImage, Function, Array } kind = None;
	uint32_t elem = 0;  // vector/array/pointer/sampledimage: element / pointee / image type id
	uint32_t count = 0; // vector: components; matrix: columns
	uint32_t storage = 0;
	std::vector<uint32_t> members;
	bool image2D = false;
};

struct Deco
{
	int location = -1, builtin = -1, binding = -1, set = -1, component = 0;
	bool flat = false, noPersp = false, block = false;
	std::map<uint32_t, int> memberBuiltin;
	std::map<uint32_t, uint32_t> memberOffset, memberStride; // push-constant block members: Offset, MatrixStride
	std::map<uint32_t, bool> memberRowMajor;
};

struct Value
{
	enum Kind { None, Vec, Pointer, SampledImage, IntConst } kind = None;
	int n = 0;
	swcu_shader_operand c[16]; // a matrix is held column by column: element (column j, row i) = c[j * rows + i]
	int rows = 0;              // matrix: components per column (0 = not a matrix)
	// pointer
	uint32_t var = 0;
	uint32_t pcType = 0, pcOffset = 0, pcStride = 0; // pointer into the push-constant block: pointee type, byte offset, stride between
	bool pcRowMajor = false;                         //   the columns (ColMajor) / rows (RowMajor) of the matrix it points into
	int ubo = -1;       // the block is uniform block `ubo` of the module (swcu_shader_info::uniformSet / uniformBinding), not the push constants
	int member = -1;    // struct member index (gl_PerVertex)
	int component = -1; // vector component selected by an access chain
	uint32_t ival = 0;
};

struct Fail
{
	char *err;
	size_t len;
	int operator()(const char *fmt, ...) const
	{
		if(err && len)
		{
			va_list ap;
			va_start(ap, fmt);
			vsnprintf(err, len, fmt, ap);
			va_end(ap);
		}
		return SWCU_E_UNSUPPORTED;
	}
};

} // namespace

extern "C" int swcu_shader_translate(const uint32_t *code, uint32_t words, swcu_shader_info *out, char *err, size_t errlen)
{
	Fail fail{ err, errlen };
	if(err && errlen) err[0] = 0;
	if(!code || !out || words < 5) return fail("null or truncated module"), SWCU_E_INVALID;
	if(code[0] != 0x07230203u) return fail("bad SPIR-V magic 0x%08x", code[0]), SWCU_E_INVALID;
	memset(out, 0, sizeof(*out));

	const uint32_t bound = code[3];
	if(bound == 0 || bound > (1u << 20)) return fail("unreasonable id bound %u", bound), SWCU_E_INVALID;
	std::vector<Type> types(bound);
	std::vector<Deco> decos(bound);
	std::vector<Value> values(bound);
	std::vector<uint32_t> varType(bound, 0); // pointer type id of OpVariable
	uint32_t entry = 0;
	int model = -1;
	bool inFunction = false, sawLabel = false, done = false;
	int functions = 0;
	bool sampled = false, posWritten = false;

	auto okId = [&](uint32_t id) { return id > 0 && id < bound; };

	for(uint32_t pc = 5; pc < words;)
	{
		const uint32_t w0 = code[pc];
		const uint32_t len = w0 >> 16, op = w0 & 0xFFFF;
		if(len == 0 || pc + len > words) return fail("malformed instruction at word %u", pc), SWCU_E_INVALID;
		const uint32_t *a = code + pc + 1; // operands
		const uint32_t na = len - 1;
		pc += len;
#define NEED(k) do { if(na < (k)) return fail("opcode %u: too few operands", op), SWCU_E_INVALID; } while(0)
#define ID(x) do { if(!okId(x)) return fail("opcode %u: id %u out of range", op, (unsigned)(x)), SWCU_E_INVALID; } while(0)
		switch(op)
		{
		case OpNop: case OpSource: case OpSourceContinued: case OpSourceExtension: case OpName: case OpMemberName: case OpString:
		case OpLine: case OpNoLine: case OpModuleProcessed: case OpExtInstImport:
			break;
		case OpExtension:
			break; // declarative; any *use* of an extension feature is rejected by opcode below
		case OpCapability:
			NEED(1);
			if(a[0] != 1 /*Shader*/ && a[0] != 0 /*Matrix (implied)*/) return fail("capability %u outside the subset", a[0]);
			break;
		case OpMemoryModel:
			NEED(2);
			if(a[0] != 0 /*Logical*/) return fail("addressing model %u unsupported", a[0]);
			break;
		case OpEntryPoint:
			NEED(2);
			if(entry) return fail("more than one entry point");
			if(a[0] != 0 && a[0] != 4) return fail("execution model %u outside the subset (vertex, fragment)", a[0]);
			model = (int)a[0];
			entry = a[1];
			ID(entry);
			break;
		case OpExecutionMode:
			NEED(2);
			if(a[1] != 7 /*OriginUpperLeft*/) return fail("execution mode %u outside the subset", a[1]);
			break;
		case OpDecorate:
		{
			NEED(2);
			ID(a[0]);
			Deco &d = decos[a[0]];
			switch(a[1])
			{
			case DecLocation: NEED(3); d.location = (int)a[2]; break;
			case DecBuiltIn: NEED(3); d.builtin = (int)a[2]; break;
			case DecBinding: NEED(3); d.binding = (int)a[2]; break;
			case DecDescriptorSet: NEED(3); d.set = (int)a[2]; break;
			case DecComponent: NEED(3); d.component = (int)a[2]; break;
			case DecFlat: d.flat = true; break;
			case DecNoPerspective: d.noPersp = true; break;
			case DecBlock: d.block = true; break;
			case DecRelaxedPrecision: break;
			default: return fail("decoration %u outside the subset", a[1]);
			}
			break;
		}
		case OpMemberDecorate:
			NEED(3);
			ID(a[0]);
			if(a[2] == DecBuiltIn) { NEED(4); decos[a[0]].memberBuiltin[a[1]] = (int)a[3]; }
			else if(a[2] == DecOffset) { NEED(4); decos[a[0]].memberOffset[a[1]] = a[3]; }
			else if(a[2] == DecMatrixStride) { NEED(4); decos[a[0]].memberStride[a[1]] = a[3]; }
			else if(a[2] == DecColMajor) decos[a[0]].memberRowMajor[a[1]] = false;
			else if(a[2] == DecRowMajor) decos[a[0]].memberRowMajor[a[1]] = true;
			else if(a[2] != DecRelaxedPrecision) return fail("member decoration %u outside the subset", a[2]);
			break;
		case OpTypeVoid: NEED(1); ID(a[0]); types[a[0]].kind = Type::Void; break;
		case OpTypeFloat:
			NEED(2); ID(a[0]);
			if(a[1] != 32) return fail("float width %u unsupported", a[1]);
			types[a[0]].kind = Type::Float;
			break;
		case OpTypeInt:
			NEED(3); ID(a[0]);
			if(a[1] != 32) return fail("int width %u unsupported", a[1]);
			types[a[0]].kind = Type::Int;
			break;
		case OpTypeVector:
			NEED(3); ID(a[0]); ID(a[1]);
			if(types[a[1]].kind != Type::Float || a[2] < 2 || a[2] > 4) return fail("only float vectors of 2..4 components are supported");
			types[a[0]].kind = Type::Vector; types[a[0]].elem = a[1]; types[a[0]].count = a[2];
			break;
		case OpTypeMatrix:
			NEED(3); ID(a[0]); ID(a[1]);
			if(types[a[1]].kind != Type::Vector || a[2] < 2 || a[2] > 4) return fail("only float matrices of 2..4 columns are supported");
			types[a[0]].kind = Type::Matrix; types[a[0]].elem = a[1]; types[a[0]].count = a[2];
			break;
		case OpTypeArray:
			NEED(3); ID(a[0]); ID(a[1]);
			types[a[0]].kind = Type::Array; types[a[0]].elem = a[1]; // only legal inside gl_PerVertex (clip/cull distance), never accessed
			break;
		case OpTypeStruct:
			NEED(1); ID(a[0]);
			types[a[0]].kind = Type::Struct;
			for(uint32_t i = 1; i < na; i++) { ID(a[i]); types[a[0]].members.push_back(a[i]); }
			break;
		case OpTypePointer:
			NEED(3); ID(a[0]); ID(a[2]);
			types[a[0]].kind = Type::Pointer; types[a[0]].storage = a[1]; types[a[0]].elem = a[2];
			break;
		case OpTypeFunction: NEED(2); ID(a[0]); types[a[0]].kind = Type::Function; if(na > 2) return fail("function parameters unsupported"); break;
		case OpTypeImage:
			NEED(8); ID(a[0]); ID(a[1]);
			types[a[0]].kind = Type::Image;
			// sampled type float, Dim 2D, not depth, not arrayed, single-sampled, sampled=1
			types[a[0]].image2D = types[a[1]].kind == Type::Float && a[2] == 1 && a[3] != 1 && a[4] == 0 && a[5] == 0 && a[6] == 1;
			break;
		case OpTypeSampledImage:
			NEED(2); ID(a[0]); ID(a[1]);
			types[a[0]].kind = Type::SampledImage; types[a[0]].elem = a[1];
			break;
		case OpConstant:
		{
			NEED(3); ID(a[0]); ID(a[1]);
			Value &v = values[a[1]];
			if(types[a[0]].kind == Type::Float) { v.kind = Value::Vec; v.n = 1; v.c[0] = { SWCU_SRC_CONST, a[2] }; }
			else if(types[a[0]].kind == Type::Int) { v.kind = Value::IntConst; v.ival = a[2]; }
			else return fail("constant of unsupported type");
			break;
		}
		case OpConstantComposite:
		{
			NEED(2); ID(a[0]); ID(a[1]);
			if(types[a[0]].kind != Type::Vector) return fail("composite constant must be a float vector");
			Value &v = values[a[1]];
			v.kind = Value::Vec; v.n = 0;
			for(uint32_t i = 2; i < na; i++)
			{
				ID(a[i]);
				const Value &e = values[a[i]];
				if(e.kind != Value::Vec || e.n != 1 || v.n >= 4) return fail("bad composite constant");
				v.c[v.n++] = e.c[0];
			}
			if(v.n != (int)types[a[0]].count) return fail("composite constant arity mismatch");
			break;
		}
		case OpVariable:
		{
			NEED(3); ID(a[0]); ID(a[1]);
			if(na > 3) return fail("variable initialisers unsupported");
			if(types[a[0]].kind != Type::Pointer) return fail("variable type is not a pointer");
			if(a[2] != SCInput && a[2] != SCOutput && a[2] != SCUniformConstant && a[2] != SCPushConstant && a[2] != SCUniform) return fail("storage class %u outside the subset", a[2]);
			varType[a[1]] = a[0];
			Value &v = values[a[1]];
			v.kind = Value::Pointer; v.var = a[1];
			if(a[2] == SCPushConstant)
			{
				// the push-constant block (sw::DrawData::pushConstants, Renderer.hpp:110): a Block struct of float scalars / vectors / matrices
				if(model != 0) return fail("push constants are only supported in the vertex stage");
				if(types[types[a[0]].elem].kind != Type::Struct) return fail("push-constant variable must be a Block struct");
				v.pcType = types[a[0]].elem;
			}
			else if(a[2] == SCUniform)
			{
				// a uniform buffer (VK_DESCRIPTOR_TYPE_UNIFORM_BUFFER): a Block struct of float scalars / vectors / matrices behind the
				// BufferDescriptor of (DescriptorSet, Binding); addressed like the push-constant block
				if(model != 0) return fail("uniform buffers are only supported in the vertex stage");
				const uint32_t st = types[a[0]].elem;
				if(types[st].kind != Type::Struct || !decos[st].block) return fail("uniform variable must be a Block struct");
				const Deco &vd = decos[a[1]];
				if(vd.set < 0 || vd.binding < 0) return fail("uniform buffer without DescriptorSet/Binding");
				if(out->uniformCount >= SWCU_MAX_UNIFORM_BUFFERS) return fail("more than %d uniform buffers", SWCU_MAX_UNIFORM_BUFFERS);
				v.ubo = (int)out->uniformCount;
				out->uniformSet[out->uniformCount] = (uint32_t)vd.set;
				out->uniformBinding[out->uniformCount] = (uint32_t)vd.binding;
				out->uniformCount++;
				v.pcType = st;
			}
			break;
		}
		case OpFunction:
			NEED(4);
			if(++functions > 1) return fail("only a single (inlined) function is supported");
			if(a[1] != entry) return fail("function is not the entry point");
			inFunction = true;
			break;
		case OpLabel:
			if(!inFunction) return fail("label outside a function"), SWCU_E_INVALID;
			if(sawLabel) return fail("control flow (more than one block) outside the subset");
			sawLabel = true;
			break;
		case OpReturn: done = true; break;
		case OpFunctionEnd: inFunction = false; break;

		case OpAccessChain: case OpInBoundsAccessChain:
		{
			NEED(4); ID(a[0]); ID(a[1]); ID(a[2]);
			const Value &base = values[a[2]];
			if(base.kind == Value::Pointer && base.pcType)
			{
				// into the push-constant block or a uniform block: member [, column [, row]] / member [, component], every index a constant
				Value v = base;
				for(uint32_t i = 3; i < na; i++)
				{
					ID(a[i]);
					if(values[a[i]].kind != Value::IntConst) return fail("access chain index into a push-constant / uniform block must be an integer constant");
					const uint32_t idx = values[a[i]].ival;
					const Type &t = types[v.pcType];
					if(t.kind == Type::Struct)
					{
						if(idx >= t.members.size()) return fail("block member index out of range");
						const Deco &sd = decos[v.pcType];
						auto off = sd.memberOffset.find(idx);
						if(off == sd.memberOffset.end()) return fail("block member without Offset");
						v.pcOffset += off->second;
						v.pcType = t.members[idx];
						if(types[v.pcType].kind == Type::Matrix)
						{
							auto st = sd.memberStride.find(idx);
							if(st == sd.memberStride.end()) return fail("matrix member without MatrixStride");
							auto rm = sd.memberRowMajor.find(idx);
							v.pcStride = st->second;
							v.pcRowMajor = rm != sd.memberRowMajor.end() && rm->second;
						}
					}
					else if(t.kind == Type::Matrix)
					{
						if(idx >= t.count) return fail("matrix column index out of range");
						v.pcOffset += v.pcRowMajor ? 4 * idx : v.pcStride * idx; // column idx: its rows lie pcStride apart when RowMajor
						v.pcType = t.elem;
						if(!v.pcRowMajor) v.pcStride = 4;
					}
					else if(t.kind == Type::Vector)
					{
						if(idx >= t.count) return fail("component index out of range");
						v.pcOffset += (v.pcStride ? v.pcStride : 4) * idx;
						v.pcType = t.elem;
					}
					else return fail("access chain into an unsupported type of a push-constant / uniform block");
				}
				values[a[1]] = v;
				break;
			}
			if(base.kind != Value::Pointer || base.member >= 0 || base.component >= 0) return fail("access chain on an unsupported base");
			if(na != 4) return fail("only single-index access chains are supported");
			ID(a[3]);
			if(values[a[3]].kind != Value::IntConst) return fail("access chain index must be an integer constant");
			const Type &pt = types[varType[base.var]];
			const Type &obj = types[pt.elem];
			Value v = base;
			if(obj.kind == Type::Struct) v.member = (int)values[a[3]].ival;
			else if(obj.kind == Type::Vector)
			{
				if(values[a[3]].ival >= obj.count) return fail("component index out of range");
				v.component = (int)values[a[3]].ival;
			}
			else return fail("access chain into an unsupported type");
			values[a[1]] = v;
			break;
		}
		case OpLoad:
		{
			NEED(3); ID(a[0]); ID(a[1]); ID(a[2]);
			if(na > 3) return fail("memory access operands unsupported");
			const Value &p = values[a[2]];
			if(p.kind != Value::Pointer) return fail("load from a non-pointer");
			if(p.pcType)
			{
				// scalar / vector / matrix of the push-constant block: one SWCU_SRC_PUSH operand per 32-bit word
				const Type &t = types[p.pcType];
				Value v; v.kind = Value::Vec;
				const int ubo = p.ubo;
				auto word = [&](uint32_t byteOffset) -> swcu_shader_operand {
					if(ubo >= 0) return { SWCU_SRC_UNIFORM, ((uint32_t)ubo << 16) | (byteOffset / 4) };
					return { SWCU_SRC_PUSH, byteOffset / 4 };
				};
				bool bad = false;
				// (a uniform block: words below 64 KiB, maxUniformBufferRange of the reference is 65536 bytes)
				auto check = [&](uint32_t byteOffset) { if((byteOffset & 3) || byteOffset / 4 >= (ubo >= 0 ? 16384u : (uint32_t)SWCU_MAX_PUSH_WORDS)) bad = true; return byteOffset; };
				if(t.kind == Type::Float) { v.n = 1; v.c[0] = word(check(p.pcOffset)); }
				else if(t.kind == Type::Vector)
				{
					v.n = (int)t.count;
					for(int i = 0; i < v.n; i++) v.c[i] = word(check(p.pcOffset + (p.pcStride ? p.pcStride : 4) * i));
				}
				else if(t.kind == Type::Matrix)
				{
					const int rows = (int)types[t.elem].count, cols = (int)t.count;
					v.n = rows * cols; v.rows = rows;
					for(int j = 0; j < cols; j++)
						for(int i = 0; i < rows; i++)
							v.c[j * rows + i] = word(check(p.pcOffset + (p.pcRowMajor ? p.pcStride * i + 4 * j : p.pcStride * j + 4 * i)));
				}
				else return fail("load of an unsupported type from a push-constant / uniform block");
				if(bad) return fail("%s access beyond %d bytes or unaligned", ubo >= 0 ? "uniform-buffer" : "push-constant", ubo >= 0 ? 65536 : 4 * SWCU_MAX_PUSH_WORDS);
				values[a[1]] = v;
				break;
			}
			const Type &pt = types[varType[p.var]];
			const Type &obj = types[pt.elem];
			const Deco &d = decos[p.var];
			Value v;
			if(pt.storage == SCUniformConstant)
			{
				if(model != 4) return fail("sampled images are only supported in the fragment stage");
				if(obj.kind != Type::SampledImage || !types[obj.elem].image2D) return fail("only a combined 2D float image sampler is supported");
				if(d.set < 0 || d.binding < 0) return fail("sampled image without DescriptorSet/Binding");
				v.kind = Value::SampledImage; v.var = p.var;
			}
			else if(pt.storage == SCInput)
			{
				if(d.builtin >= 0) return fail("built-in input %d outside the subset", d.builtin);
				// inputMask / flatMask / noPerspectiveMask are 32-bit (bit = location * 4 + component): vertex inputs at locations 0..7,
				// fragment inputs at the locations the vertex stage can write (SWCU_MAX_VARYING_COMPONENTS / 4)
				const int maxLoc = model == 4 ? SWCU_MAX_VARYING_COMPONENTS / 4 : 8;
				if(d.location < 0 || d.location >= maxLoc || d.location >= SWCU_MAX_INPUTS) return fail("input without a supported Location (0..%d)", maxLoc - 1);
				if(d.component != 0) return fail("Component decoration unsupported");
				int n = obj.kind == Type::Float ? 1 : obj.kind == Type::Vector ? (int)obj.count : 0;
				if(!n) return fail("input of unsupported type");
				v.kind = Value::Vec;
				if(p.component >= 0) { v.n = 1; v.c[0] = { SWCU_SRC_INPUT, (uint32_t)(d.location * 4 + p.component) }; }
				else { v.n = n; for(int i = 0; i < n; i++) v.c[i] = { SWCU_SRC_INPUT, (uint32_t)(d.location * 4 + i) }; }
			}
			else return fail("load from an output variable unsupported");
			values[a[1]] = v;
			break;
		}
		case OpCopyObject:
			NEED(3); ID(a[1]); ID(a[2]);
			values[a[1]] = values[a[2]];
			break;
		case OpCompositeConstruct:
		{
			NEED(2); ID(a[0]); ID(a[1]);
			if(types[a[0]].kind != Type::Vector) return fail("composite construct of a non-vector");
			Value v; v.kind = Value::Vec;
			for(uint32_t i = 2; i < na; i++)
			{
				ID(a[i]);
				const Value &e = values[a[i]];
				if(e.kind != Value::Vec || e.rows) return fail("composite construct from a non-value");
				for(int k = 0; k < e.n; k++) { if(v.n >= 4) return fail("vector too long"); v.c[v.n++] = e.c[k]; }
			}
			if(v.n != (int)types[a[0]].count) return fail("composite construct arity mismatch");
			values[a[1]] = v;
			break;
		}
		case OpCompositeExtract:
		{
			NEED(4); ID(a[1]); ID(a[2]);
			const Value &s = values[a[2]];
			if(s.kind == Value::Vec && s.rows) // column (and row) of a matrix
			{
				const uint32_t cols = (uint32_t)(s.n / s.rows);
				if((na != 4 && na != 5) || a[3] >= cols || (na == 5 && a[4] >= (uint32_t)s.rows)) return fail("unsupported matrix extract");
				Value v; v.kind = Value::Vec;
				if(na == 5) { v.n = 1; v.c[0] = s.c[a[3] * s.rows + a[4]]; }
				else { v.n = s.rows; for(int i = 0; i < s.rows; i++) v.c[i] = s.c[a[3] * s.rows + i]; }
				values[a[1]] = v;
				break;
			}
			if(s.kind != Value::Vec || na != 4 || a[3] >= (uint32_t)s.n) return fail("unsupported composite extract");
			Value v; v.kind = Value::Vec; v.n = 1; v.c[0] = s.c[a[3]];
			values[a[1]] = v;
			break;
		}
		case OpCompositeInsert:
		{
			NEED(5); ID(a[1]); ID(a[2]); ID(a[3]);
			const Value &obj = values[a[2]];
			Value v = values[a[3]];
			if(obj.kind != Value::Vec || obj.n != 1 || v.kind != Value::Vec || v.rows || na != 5 || a[4] >= (uint32_t)v.n) return fail("unsupported composite insert");
			v.c[a[4]] = obj.c[0];
			values[a[1]] = v;
			break;
		}
		case OpVectorShuffle:
		{
			NEED(4); ID(a[1]); ID(a[2]); ID(a[3]);
			const Value &x = values[a[2]], &y = values[a[3]];
			if(x.kind != Value::Vec || y.kind != Value::Vec || x.rows || y.rows || na - 4 > 4 || na - 4 < 2) return fail("unsupported vector shuffle");
			Value v; v.kind = Value::Vec;
			for(uint32_t i = 4; i < na; i++)
			{
				uint32_t s = a[i];
				if(s == 0xFFFFFFFFu) v.c[v.n++] = { SWCU_SRC_CONST, 0 };
				else if(s < (uint32_t)x.n) v.c[v.n++] = x.c[s];
				else if(s < (uint32_t)(x.n + y.n)) v.c[v.n++] = y.c[s - x.n];
				else return fail("shuffle index out of range");
			}
			values[a[1]] = v;
			break;
		}
		case OpFNegate: case OpFAdd: case OpFSub: case OpFMul: case OpVectorTimesScalar: case OpMatrixTimesScalar:
		case OpMatrixTimesVector: case OpVectorTimesMatrix: case OpDot:
		{
			// straight-line float arithmetic of the vertex stage -> scalar program steps (SpirvShaderArithmetic.cpp)
			NEED(3); ID(a[0]); ID(a[1]); ID(a[2]);
			if(model != 0) return fail("arithmetic in the fragment stage is outside the subset (opcode %u)", op);
			const bool unary = op == OpFNegate;
			if(!unary) { NEED(4); ID(a[3]); }
			const Value &x = values[a[2]];
			const Value &y = unary ? x : values[a[3]];
			if(x.kind != Value::Vec || y.kind != Value::Vec) return fail("arithmetic on a non-value (opcode %u)", op);
			bool full = false;
			auto emit = [&](uint32_t o, swcu_shader_operand p0, swcu_shader_operand p1, swcu_shader_operand p2) -> swcu_shader_operand {
				if(out->programLength >= SWCU_MAX_PROGRAM) { full = true; return { SWCU_SRC_CONST, 0 }; }
				swcu_shader_op &st = out->program[out->programLength];
				st.op = o; st.a = p0; st.b = p1; st.c = p2;
				return { SWCU_SRC_TEMP, out->programLength++ };
			};
			const swcu_shader_operand none = { SWCU_SRC_CONST, 0 };
			Value v; v.kind = Value::Vec;
			switch(op)
			{
			case OpFNegate:
				if(x.rows) return fail("matrix negate outside the subset");
				v.n = x.n;
				for(int i = 0; i < x.n; i++) v.c[i] = emit(SWCU_OP_NEG, x.c[i], none, none);
				break;
			case OpFAdd: case OpFSub: case OpFMul:
				if(x.rows || y.rows || x.n != y.n || x.n > 4) return fail("component-wise arithmetic on mismatched or matrix operands");
				v.n = x.n;
				for(int i = 0; i < x.n; i++) v.c[i] = emit(op == OpFAdd ? SWCU_OP_ADD : op == OpFSub ? SWCU_OP_SUB : SWCU_OP_MUL, x.c[i], y.c[i], none);
				break;
			case OpVectorTimesScalar: case OpMatrixTimesScalar: // :57-75 pattern: lhs.Float(i) * rhs.Float(0)
				if(y.n != 1 || (op == OpVectorTimesScalar && x.rows) || (op == OpMatrixTimesScalar && !x.rows)) return fail("bad operands of a times-scalar product");
				v.n = x.n; v.rows = x.rows;
				for(int i = 0; i < x.n; i++) v.c[i] = emit(SWCU_OP_MUL, x.c[i], y.c[0], none);
				break;
			case OpMatrixTimesVector: // :39-55: v_i = M[i,0] * x_0, then v_i = MulAdd(M[i,j], x_j, v_i)
			{
				if(!x.rows || y.rows || y.n != x.n / x.rows) return fail("bad operands of OpMatrixTimesVector");
				v.n = x.rows;
				for(int i = 0; i < x.rows; i++)
				{
					swcu_shader_operand acc = emit(SWCU_OP_MUL, x.c[i], y.c[0], none);
					for(int j = 1; j < y.n; j++) acc = emit(SWCU_OP_FMA, x.c[i + x.rows * j], y.c[j], acc);
					v.c[i] = acc;
				}
				break;
			}
			case OpVectorTimesMatrix: // :57-73: v_i = x_0 * M[0,i], then v_i = MulAdd(x_j, M[j,i], v_i)
			{
				if(x.rows || !y.rows || x.n != y.rows) return fail("bad operands of OpVectorTimesMatrix");
				v.n = y.n / y.rows;
				for(int i = 0; i < v.n; i++)
				{
					swcu_shader_operand acc = emit(SWCU_OP_MUL, x.c[0], y.c[i * y.rows], none);
					for(int j = 1; j < x.n; j++) acc = emit(SWCU_OP_FMA, x.c[j], y.c[i * y.rows + j], acc);
					v.c[i] = acc;
				}
				break;
			}
			default: // OpDot, :611-621: d = x_0 * y_0, then d = MulAdd(x_i, y_i, d)
			{
				if(x.rows || y.rows || x.n != y.n || x.n < 2 || x.n > 4) return fail("bad operands of OpDot");
				swcu_shader_operand acc = emit(SWCU_OP_MUL, x.c[0], y.c[0], none);
				for(int i = 1; i < x.n; i++) acc = emit(SWCU_OP_FMA, x.c[i], y.c[i], acc);
				v.n = 1; v.c[0] = acc;
				break;
			}
			}
			if(full) return fail("vertex shader arithmetic exceeds %d scalar steps", SWCU_MAX_PROGRAM);
			values[a[1]] = v;
			break;
		}
		case OpImageSampleImplicitLod:
		{
			NEED(4); ID(a[1]); ID(a[2]); ID(a[3]);
			if(na > 4) return fail("image operands (bias/offset/...) outside the subset");
			if(model != 4) return fail("implicit-LOD sampling outside the fragment stage");
			if(sampled) return fail("more than one image sample outside the subset");
			const Value &img = values[a[2]], &uv = values[a[3]];
			if(img.kind != Value::SampledImage) return fail("sample of a non-sampled-image");
			if(uv.kind != Value::Vec || uv.n < 2) return fail("sample coordinate must have 2 components");
			for(int k = 0; k < 2; k++)
				if(uv.c[k].kind == SWCU_SRC_TEXEL) return fail("dependent texture reads outside the subset");
			sampled = true;
			out->usesTexture = 1;
			out->textureSet = (uint32_t)decos[img.var].set;
			out->textureBinding = (uint32_t)decos[img.var].binding;
			out->texCoord[0] = uv.c[0];
			out->texCoord[1] = uv.c[1];
			Value v; v.kind = Value::Vec; v.n = 4;
			for(uint32_t k = 0; k < 4; k++) v.c[k] = { SWCU_SRC_TEXEL, k };
			values[a[1]] = v;
			break;
		}
		case OpStore:
		{
			NEED(2); ID(a[0]); ID(a[1]);
			if(na > 2) return fail("memory access operands unsupported");
			const Value &p = values[a[0]], &v = values[a[1]];
			if(p.kind != Value::Pointer || v.kind != Value::Vec || v.rows || p.pcType) return fail("unsupported store");
			const Type &pt = types[varType[p.var]];
			if(pt.storage != SCOutput) return fail("store to a non-output variable");
			const Type &obj = types[pt.elem];
			const Deco &d = decos[p.var];
			int builtin = d.builtin;
			if(obj.kind == Type::Struct)
			{
				if(p.member < 0) return fail("whole-struct store unsupported");
				auto it = decos[pt.elem].memberBuiltin.find((uint32_t)p.member);
				if(it == decos[pt.elem].memberBuiltin.end()) return fail("store to an undecorated block member");
				builtin = it->second;
			}
			if(builtin >= 0)
			{
				if(model != 0) return fail("built-in output in the fragment stage outside the subset (FragDepth etc.)");
				if(builtin == BuiltInPointSize) // read by point draws only (DrawCall::setupPoint, Renderer.cpp:1151)
				{
					if(v.n != 1 || v.c[0].kind == SWCU_SRC_TEXEL) return fail("gl_PointSize must be stored as a float");
					out->pointSize = v.c[0];
					out->writesPointSize = 1;
					break;
				}
				if(builtin != BuiltInPosition) return fail("built-in output %d outside the subset", builtin);
				if(v.n != 4) return fail("gl_Position must be stored as a vec4");
				for(int k = 0; k < 4; k++) out->position[k] = v.c[k];
				posWritten = true;
				break;
			}
			if(d.location < 0) return fail("output without Location");
			if(d.component != 0) return fail("Component decoration unsupported");
			if(model == 4 && d.location != 0) return fail("only colour attachment 0 is supported");
			if(d.location * 4 + 4 > SWCU_MAX_VARYING_COMPONENTS) return fail("output location %d beyond the supported range", d.location);
			if(p.component >= 0)
			{
				if(v.n != 1) return fail("component store arity mismatch");
				out->output[d.location * 4 + p.component] = v.c[0];
				out->outputMask |= 1u << (d.location * 4 + p.component);
			}
			else
			{
				int n = obj.kind == Type::Float ? 1 : obj.kind == Type::Vector ? (int)obj.count : 0;
				if(n == 0 || v.n != n) return fail("output store arity mismatch");
				for(int k = 0; k < n; k++)
				{
					out->output[d.location * 4 + k] = v.c[k];
					out->outputMask |= 1u << (d.location * 4 + k);
				}
			}
			break;
		}
		default:
			return fail("opcode %u outside the supported subset", op);
		}
#undef NEED
#undef ID
	}
	if(!entry || model < 0) return fail("no entry point");
	if(!done) return fail("entry point has no OpReturn");
	out->stage = (uint32_t)model;

	// inputs actually consumed by the results
	auto use = [&](const swcu_shader_operand &o) { if(o.kind == SWCU_SRC_INPUT && o.value < 32) out->inputMask |= 1u << o.value; };
	if(model == 0)
	{
		// a VS that never writes gl_Position is legal (tests/VulkanUnitTests/DrawTests.cpp:26-76) but outside the subset
		if(!posWritten) return fail("vertex shader does not write gl_Position");
		for(int k = 0; k < 4; k++)
		{
			if(out->position[k].kind == SWCU_SRC_TEXEL) return fail("bad position operand");
			use(out->position[k]);
		}
		for(int k = 0; k < SWCU_MAX_VARYING_COMPONENTS; k++)
			if(out->outputMask >> k & 1) use(out->output[k]);
		for(uint32_t i = 0; i < out->programLength; i++) { use(out->program[i].a); use(out->program[i].b); use(out->program[i].c); }
		if(out->writesPointSize) use(out->pointSize);
	}
	else
	{
		if(!(out->outputMask & 0xF)) return fail("fragment shader does not write colour output 0");
		out->outputMask &= 0xF;
		for(int k = 0; k < 4; k++)
			if(out->outputMask >> k & 1) use(out->output[k]);
		if(out->usesTexture) { use(out->texCoord[0]); use(out->texCoord[1]); }
		if(out->inputMask >> SWCU_MAX_VARYING_COMPONENTS) return fail("fragment input location beyond the supported range");
		// interpolation qualifiers of the consumed inputs
		for(uint32_t id = 1; id < bound; id++)
		{
			if(!varType[id] || types[varType[id]].storage != SCInput || decos[id].location < 0) continue;
			for(int c = 0; c < 4; c++)
			{
				if(decos[id].location * 4 + c >= 32) continue;
				uint32_t bit = 1u << (decos[id].location * 4 + c);
				if(!(out->inputMask & bit)) continue;
				if(decos[id].flat) out->flatMask |= bit;
				if(decos[id].noPersp) out->noPerspectiveMask |= bit;
			}
		}
	}
	return SWCU_OK;
}
